#!/bin/bash
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/gpus.txt
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/bench_v6_n1.json 2> gpurun_out/bench_v6_n1.err; cat gpurun_out/bench_v6_n1.json
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/bench_v6_n2.json 2> gpurun_out/bench_v6_n2.err; tail -5 gpurun_out/bench_v6_n2.err; cat gpurun_out/bench_v6_n2.json
timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu --batch 32768 > gpurun_out/bench_v6_b32k.json 2> gpurun_out/bench_v6_b32k.err; cat gpurun_out/bench_v6_b32k.json
