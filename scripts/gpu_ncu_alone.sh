#!/bin/bash
# ncu source-level capture of one straggler instance solved alone; $1 = library suffix (default product library)
mkdir -p gpurun_out
lib=/root/repo/mpc_collisionavoidance_b200/libusvmpc.so
[ -n "$1" ] && lib=/root/repo/mpc_collisionavoidance_b200/libusvmpc_$1.so
cd scripts
USVMPC_LIB=$lib timeout 900 ncu --set full --import-source on --clock-control none -k regex:nmpc_solve -c 1 -f -o ../gpurun_out/alone_${1:-base} python gpu_alone.py 2>&1 | tail -5
