#!/bin/bash
# the three multi-GPU lines that matter (weak, strong, config 5) with few steps: N=$1
N=${1:-8}
mkdir -p gpurun_out
run() { name=$1; shift; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N "$@" > gpurun_out/bench_$name.json 2> gpurun_out/bench_$name.err || tail -5 gpurun_out/bench_$name.err; python3 -c "
import json
d=json.loads(open('gpurun_out/bench_$name.json').read().strip().split('\n')[-1]); print('$name', d['n_gpus'], d['scaling'], d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['config']['global_batch'])"; }
run n${N}_weak --steps 3 --warmup 3
run n${N}_strong --steps 3 --warmup 3 --strong
run n${N}_cfg5 --config 5 --steps 2 --warmup 3
