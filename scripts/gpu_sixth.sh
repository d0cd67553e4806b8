#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -2 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/bench_v11.json 2> gpurun_out/bench_v11.err; cat gpurun_out/bench_v11.json
timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu --batch 32768 > gpurun_out/bench_v11_b32k.json 2> gpurun_out/bench_v11_b32k.err; cat gpurun_out/bench_v11_b32k.json
timeout 900 ncu --set full --clock-control none --import-source on -k regex:nmpc_solve -s 3 -c 1 -f -o gpurun_out/prof_v11 python bench.py --steps 1 --warmup 3 --no-cpu --batch 512 > gpurun_out/ncu_full_v11.log 2>&1
tail -2 gpurun_out/ncu_full_v11.log
