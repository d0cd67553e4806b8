#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for h in 0 1; do
  USVMPC_HELPERS_OPT=$h timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/h$h.json 2>gpurun_out/h$h.err
  python3 -c "
import json
d=json.loads(open('gpurun_out/h$h.json').read().strip().split('\n')[-1]); print('helpers=$h B=4096', d['value'], d['ms_per_step'], d['e2e']['value'])"
  USVMPC_HELPERS_OPT=$h timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu --batch 32768 > gpurun_out/h${h}_b32k.json 2>/dev/null
  python3 -c "
import json
d=json.loads(open('gpurun_out/h${h}_b32k.json').read().strip().split('\n')[-1]); print('helpers=$h B=32768', d['value'], d['ms_per_step'])"
done
USVMPC_LIB=/root/repo/mpc_collisionavoidance_b200/libusvmpc_ph1.so timeout 300 python scripts/gpu_phase_profile.py 2>&1 | grep -v PROFA | head -8
