#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/bench_v8.json 2> gpurun_out/bench_v8.err; tail -3 gpurun_out/bench_v8.err; cat gpurun_out/bench_v8.json
timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu --batch 32768 > gpurun_out/bench_v8_b32k.json 2> gpurun_out/bench_v8_b32k.err; cat gpurun_out/bench_v8_b32k.json
for v in 5 6; do
  USVMPC_LIB=$PWD/mpc_collisionavoidance_b200/libusvmpc_c$v.so timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu --batch 32768 > gpurun_out/bench_v8_b32k_c$v.json 2> gpurun_out/bench_v8_b32k_c$v.err
  cat gpurun_out/bench_v8_b32k_c$v.json
done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:nmpc_solve -s 3 -c 1 -f -o gpurun_out/prof_v8 python bench.py --steps 1 --warmup 3 --no-cpu --batch 512 > gpurun_out/ncu_full_v8.log 2>&1
tail -2 gpurun_out/ncu_full_v8.log
