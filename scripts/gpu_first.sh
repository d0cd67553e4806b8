#!/bin/bash
# first GPU pass: parity tests, a short bench, occupancy variants
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -30 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/bench_first.json 2> gpurun_out/bench_first.err; tail -5 gpurun_out/bench_first.err; cat gpurun_out/bench_first.json
for v in 2 3 5 7; do
  USVMPC_LIB=$PWD/mpc_collisionavoidance_b200/libusvmpc_c$v.so timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/bench_c$v.json 2> gpurun_out/bench_c$v.err
  cat gpurun_out/bench_c$v.json
done
