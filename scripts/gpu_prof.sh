#!/bin/bash
# phase clocks of the slowest instances (+ optional ncu source-level capture of one straggler solved alone: NCU=1)
mkdir -p gpurun_out
cd scripts
USVMPC_LIB=/root/repo/mpc_collisionavoidance_b200/libusvmpc_p1.so timeout 600 python gpu_phase_profile.py 2>&1 | grep -v "^$" | tail -30 | tee ../gpurun_out/phase_clocks.txt
cd ..
[ -n "$NCU" ] && bash scripts/gpu_ncu_alone.sh
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err; tail -3 gpurun_out/bench_quick.err; python3 -c "
import json
d=json.loads(open('gpurun_out/bench_quick.json').read().strip().split('\n')[-1]); print('b4096', d['value'], d['ms_per_step'], d['e2e']['value'])"
timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu --batch 32768 > gpurun_out/bench_quick_b32k.json 2>/dev/null; python3 -c "
import json
d=json.loads(open('gpurun_out/bench_quick_b32k.json').read().strip().split('\n')[-1]); print('b32768', d['value'], d['ms_per_step'])"
