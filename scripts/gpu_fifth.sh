#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -2 gpurun_out/pytest_gpu.log
for v in tma ldgsts; do
  L=$PWD/mpc_collisionavoidance_b200/libusvmpc.so; [ $v = ldgsts ] && L=$PWD/mpc_collisionavoidance_b200/libusvmpc_ldgsts.so
  USVMPC_LIB=$L timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/bench_v9_$v.json 2> gpurun_out/bench_v9_$v.err; cat gpurun_out/bench_v9_$v.json
  USVMPC_LIB=$L timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu --batch 32768 > gpurun_out/bench_v9_b32k_$v.json 2> gpurun_out/bench_v9_b32k_$v.err; cat gpurun_out/bench_v9_b32k_$v.json
done
