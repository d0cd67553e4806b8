#!/bin/bash
# multi-GPU bench lines on an N-GPU box: N=$1 (2, 4 or 8).  Weak scaling of the headline config, strong scaling of the
# headline batch, config 5 (131072 instances over 8 GPUs = 16384 per GPU) and the reference arm's launch path under torchrun.
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader > gpurun_out/gpus_n$N.txt
run() { name=$1; shift; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N "$@" > gpurun_out/bench_$name.json 2> gpurun_out/bench_$name.err || tail -5 gpurun_out/bench_$name.err; python3 -c "
import json
d=json.loads(open('gpurun_out/bench_$name.json').read().strip().split('\n')[-1]); print('$name', d['n_gpus'], d['scaling'], d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['config']['global_batch'])"; }
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/bench_n1_ref.json 2>/dev/null; python3 -c "
import json
d=json.loads(open('gpurun_out/bench_n1_ref.json').read().strip().split('\n')[-1]); print('n1', d['value'], d['ms_per_step'])"
run n${N}_weak --steps 5 --warmup 3
run n${N}_strong --steps 5 --warmup 3 --strong
run n${N}_cfg5 --config 5 --steps 3 --warmup 3
run n${N}_cfg3 --config 3 --steps 3 --warmup 3
