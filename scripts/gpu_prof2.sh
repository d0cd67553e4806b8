#!/bin/bash
# phase clocks of the straggler with two profile builds (p1 = product options, p0 = a variant to compare)
mkdir -p gpurun_out
cd scripts
for v in p1 p0; do
echo "== $v"
USVMPC_LIB=/root/repo/mpc_collisionavoidance_b200/libusvmpc_$v.so timeout 600 python gpu_phase_profile.py 2>&1 | grep -v "^$" | tail -24 | tee ../gpurun_out/phase_clocks_$v.txt
done
