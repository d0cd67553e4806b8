#!/bin/bash
# phase clocks of the straggler only (profile build p1)
mkdir -p gpurun_out
cd scripts
USVMPC_LIB=/root/repo/mpc_collisionavoidance_b200/libusvmpc_p1.so timeout 600 python gpu_phase_profile.py 2>&1 | grep -v "^$" | tail -16 | tee ../gpurun_out/phase_clocks_p1.txt
