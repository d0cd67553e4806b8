#!/bin/bash
# usage: gpu_prof_variants.sh <names...>: names starting with "np" are benchmarked (no profile clocks), others run the phase profile
mkdir -p gpurun_out
for v in "$@"; do
  lib=/root/repo/mpc_collisionavoidance_b200/libusvmpc_$v.so
  case $v in
    np*|base) [ "$v" = base ] && lib=/root/repo/mpc_collisionavoidance_b200/libusvmpc.so
       USVMPC_LIB=$lib timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/var_$v.json 2>gpurun_out/var_$v.err
       python3 -c "
import json
d=json.loads(open('gpurun_out/var_$v.json').read().strip().split('\n')[-1]); print('$v', 'B=4096', d['value'], d['ms_per_step'], d['e2e']['value'])"
       if [ -n "$BIG" ]; then
       USVMPC_LIB=$lib timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu --batch 32768 > gpurun_out/var_${v}_b32k.json 2>/dev/null
       python3 -c "
import json
d=json.loads(open('gpurun_out/var_${v}_b32k.json').read().strip().split('\n')[-1]); print('$v', 'B=32768', d['value'], d['ms_per_step'])"
       fi ;;
    *) echo "== $v"; USVMPC_LIB=$lib timeout 300 python scripts/gpu_phase_profile.py 2>&1 | grep -A1 "inst 0 sqp 100 qp 1802\|^alone\|^full\|^148" | grep -v "^--" ;;
  esac
done
