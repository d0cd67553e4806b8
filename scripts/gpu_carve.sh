#!/bin/bash
for c in 13 25 50 100; do echo "== carveout $c"; USVMPC_CARVEOUT=$c USVMPC_LIB=/root/repo/mpc_collisionavoidance_b200/libusvmpc_$1.so timeout 300 python scripts/gpu_phase_profile.py 2>&1 | grep "^full\|^alone inst 1089\|^148\|inst 0 sqp 100 qp 1802" ; done
