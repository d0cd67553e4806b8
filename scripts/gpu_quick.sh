#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err; python3 -c "
import json
d=json.loads(open('gpurun_out/bench_quick.json').read().strip().split('\n')[-1]); print(d['value'], d['ms_per_step'], d['e2e']['value'])"
timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu --batch 32768 > gpurun_out/bench_quick_b32k.json 2>/dev/null; python3 -c "
import json
d=json.loads(open('gpurun_out/bench_quick_b32k.json').read().strip().split('\n')[-1]); print(d['value'], d['ms_per_step'])"
