#!/bin/bash
# reduced round-end set (when GPU minutes are short): smoke, bench, bench --batch 32768, launch list
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 900 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; cat gpurun_out/bench_final.json | cut -c1-200
timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu --batch 32768 > gpurun_out/bench_final_b32k.json 2> gpurun_out/bench_final_b32k.err; cat gpurun_out/bench_final_b32k.json | cut -c1-200
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/launches_final.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_launch_final.log 2>&1
