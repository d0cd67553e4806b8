#!/bin/bash
# quick GPU check: parity tests, a short bench at the headline batch and at batch 32768
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | grep -E "^E  |passed|failed|^FAILED|Error" | head -40
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err; tail -3 gpurun_out/bench_quick.err; python3 -c "
import json
d=json.loads(open('gpurun_out/bench_quick.json').read().strip().split('\n')[-1]); print('b4096', d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['kernel_ms'])"
timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu --batch 32768 > gpurun_out/bench_quick_b32k.json 2>/dev/null; python3 -c "
import json
d=json.loads(open('gpurun_out/bench_quick_b32k.json').read().strip().split('\n')[-1]); print('b32768', d['value'], d['ms_per_step'])"
