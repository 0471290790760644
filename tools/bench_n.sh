#!/bin/bash
# usage: tools/bench_n.sh N [extra bench args] -- runs bench.py on N GPUs the way the driver does and prints the headline fields
N=$1; shift
if [ "$N" = 1 ]; then python bench.py --gpus 1 "$@" > gpurun_out/bench_n$N.log 2>gpurun_out/bench_n$N.err
else python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N "$@" > gpurun_out/bench_n$N.log 2>gpurun_out/bench_n$N.err; fi
grep '^{' gpurun_out/bench_n$N.log | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('N', d['n_gpus'], 'value', round(d['value']), 'ms', round(d['ms_per_step'],4), 'e2e_ms', round(d['e2e']['ms_per_step'],4), 'epoch_ms', round(d.get('e2e_epoch',{}).get('ms_per_step',0),4), 'beam', d.get('beam',{}).get('value'))"
