#!/bin/bash
# round-2 measurement pass on one B200: parity tests, bench lines of every config, the ncu launch list and the
# `ncu --set full` captures of the solve kernel at the headline batch and at batch 32768 (outputs in gpurun_out/)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | grep -E "^E  |passed|failed|^FAILED|Error" | head -40
run() { name=$1; shift; timeout 900 python bench.py "$@" > gpurun_out/bench_$name.json 2> gpurun_out/bench_$name.err || tail -3 gpurun_out/bench_$name.err; python3 -c "
import json,sys
d=json.loads(open('gpurun_out/bench_$name.json').read().strip().split('\n')[-1]); print('$name', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], 'kernel_ms', d['roofline']['kernel_ms'], d['workload_stats']['converged_frac'])"; }
run cfg2 --steps 5 --warmup 3
run cfg2_b32768 --steps 3 --warmup 3 --batch 32768 --no-cpu
run cfg3 --config 3 --steps 5 --warmup 3 --no-cpu
run cfg4 --config 4 --steps 5 --warmup 3 --no-cpu
run cfg4_fp32 --config 4 --steps 5 --warmup 3 --no-cpu --riccati-precision 32
run cfg5_1gpu --config 5 --steps 3 --warmup 3 --no-cpu
run cfg1 --config 1 --steps 20 --warmup 5 --no-cpu
if [ -n "$NCU" ]; then
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_cfg2.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_launch.log 2>&1
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:nmpc_solve_kernel -s 4 -c 1 -f -o gpurun_out/r2_cfg2_b4096 python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_b4096.log 2>&1
timeout 1500 ncu --set full --clock-control none -k regex:nmpc_solve_kernel -s 4 -c 1 -f -o gpurun_out/r2_cfg2_b32768 python bench.py --steps 2 --warmup 3 --no-cpu --batch 32768 > gpurun_out/ncu_b32768.log 2>&1
timeout 1500 ncu --set full --clock-control none -k regex:nmpc_solve_kernel -s 4 -c 1 -f -o gpurun_out/r2_cfg4_b1024 python bench.py --config 4 --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_cfg4.log 2>&1
ls -la gpurun_out/*.ncu-rep
fi
