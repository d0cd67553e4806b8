#!/bin/bash
# GPU parity tests, phase clocks of the straggler (profile build p1), then the quick bench
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | grep -E "^E  |passed|failed|^FAILED|Error" | head -20
cd scripts
USVMPC_LIB=/root/repo/mpc_collisionavoidance_b200/libusvmpc_p1.so timeout 600 python gpu_phase_profile.py 2>&1 | grep -v "^$" | tail -12 | tee ../gpurun_out/phase_clocks_p1.txt
cd ..
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err; tail -3 gpurun_out/bench_quick.err; python3 -c "
import json
d=json.loads(open('gpurun_out/bench_quick.json').read().strip().split('\n')[-1]); print('b4096', d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['kernel_ms'])"
timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu --batch 32768 > gpurun_out/bench_quick_b32k.json 2>/dev/null; python3 -c "
import json
d=json.loads(open('gpurun_out/bench_quick_b32k.json').read().strip().split('\n')[-1]); print('b32768', d['value'], d['ms_per_step'])"
