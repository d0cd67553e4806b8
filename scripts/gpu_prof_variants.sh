#!/bin/bash
# bench the geometry variants built by scripts/build_geometry_variants.py: batch 4096, batch 32768, one straggler alone
mkdir -p gpurun_out
for lib in mpc_collisionavoidance_b200/libusvmpc_t*.so; do
  n=$(basename $lib .so)
  a=$(USVMPC_LIB=/root/repo/$lib timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu 2>/dev/null | python3 -c "import json,sys; d=json.loads(sys.stdin.read().strip().split('\n')[-1]); print(d['value'], d['ms_per_step'])")
  b=$(USVMPC_LIB=/root/repo/$lib timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu --batch 32768 2>/dev/null | python3 -c "import json,sys; d=json.loads(sys.stdin.read().strip().split('\n')[-1]); print(d['value'])")
  c=$(cd scripts; USVMPC_LIB=/root/repo/$lib timeout 300 python gpu_alone.py 2>/dev/null | tail -1)
  echo "$n | b4096 $a | b32768 $b | $c" | tee -a gpurun_out/variants.txt
done
