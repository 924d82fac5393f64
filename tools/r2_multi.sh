#!/bin/bash
# Multi-GPU pass of round 2 on an N-GPU box: bench.py at N (and N/2) under torchrun, the single-process C entry orbm_allpairs_multi.
# usage: gpurun --gpus N --timeout 900 -- bash tools/r2_multi.sh N TAG
n=${1:-8}; tag=${2:-r2}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
for g in $n $((n / 2)); do
  [ $g -ge 2 ] || continue
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $g --master-addr 127.0.0.1 --master-port $((29500 + g)) bench.py --gpus $g --steps 10 --warmup 3 --no-configs \
      > gpurun_out/bench_${tag}_n$g.json 2> gpurun_out/bench_${tag}_n$g.err
  python -c "
import json
d=json.loads(open('gpurun_out/bench_${tag}_n$g.json').read().strip().splitlines()[-1])
e=d['e2e']; print('N=$g value', round(d['value']), 'e2e', round(e['value']), 'copy floor ms', e.get('copy_only_ms_per_step'), 'e2e ms', e.get('ms_per_step'), 'ratio', e.get('e2e_over_copy_floor'), 'matching', d['matching']['value'] if d.get('matching') else None)"
done
python tools/allpairs_multi_probe.py 2048 2000 2>&1 | tail -1 > gpurun_out/allpairs_multi_${tag}_n$n.json; cat gpurun_out/allpairs_multi_${tag}_n$n.json | cut -c1-900
