#!/bin/bash
set -x
mkdir -p gpurun_out
bash scripts/gpu_sanitize.sh > gpurun_out/sanitize_run.log 2>&1; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|Race reported" gpurun_out/sanitizer_*.log | head
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -2 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -1 gpurun_out/smoke.log
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_final_reference.json 2> gpurun_out/bench_final_reference.err; cat gpurun_out/bench_final_reference.json | cut -c1-400
timeout 900 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; tail -2 gpurun_out/bench_final.err; cat gpurun_out/bench_final.json
timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu --batch 32768 > gpurun_out/bench_final_b32k.json 2> gpurun_out/bench_final_b32k.err; cat gpurun_out/bench_final_b32k.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/launches_final.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_launch_final.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:nmpc_solve -s 3 -c 1 -f -o gpurun_out/prof_final python bench.py --steps 1 --warmup 3 --no-cpu --batch 512 > gpurun_out/ncu_full_final.log 2>&1
tail -2 gpurun_out/ncu_full_final.log
