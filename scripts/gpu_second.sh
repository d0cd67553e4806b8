#!/bin/bash
# second GPU pass: parity tests, bench + occupancy variants, ncu launch list and one full capture of the solve kernel
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_v2.json 2> gpurun_out/bench_v2.err; tail -3 gpurun_out/bench_v2.err; cat gpurun_out/bench_v2.json
for v in 2 3 5; do
  USVMPC_LIB=$PWD/mpc_collisionavoidance_b200/libusvmpc_c$v.so timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/bench_v2_c$v.json 2> gpurun_out/bench_v2_c$v.err
  cat gpurun_out/bench_v2_c$v.json
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_v2.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_launch.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:nmpc_solve -s 3 -c 1 -f -o gpurun_out/prof_v2 python bench.py --steps 1 --warmup 3 --no-cpu > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log
ls -la gpurun_out
