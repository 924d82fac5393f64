#!/bin/bash
# Short end-of-round pass: gpu tests, the default bench line, and a --set full re-capture of the kernels that changed after r1q.
tag=${1:-r1s}
python -m pytest tests -m gpu -x -q 2>&1 | tail -2
python bench.py --steps 20 --warmup 3 > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err
python -c "
import json
d=json.loads(open('gpurun_out/bench_$tag.json').read().strip().splitlines()[-1]); print('$tag', round(d['value']), round(d['ms_per_step'],3), round(d['e2e']['value']), d['gpu_launches'], d['roofline']['stage_ms_per_step'], d['roofline']['frac'], d['clocks'], d['matching']['value'])"
ncu --set full --clock-control none --import-source on -k regex:"k_resize|k_describe" -s 24 -c 8 \
    -f -o gpurun_out/${tag}_resize_describe python bench.py --frames 256 --unique 64 --chunk 256 --steps 2 --warmup 3 --no-cpu --no-match > gpurun_out/ncu_${tag}.log 2>&1
ls -la gpurun_out/${tag}_resize_describe.ncu-rep
