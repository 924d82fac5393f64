#!/bin/bash
# End-of-round GPU pass: the gpu test-suite, the bench line, the ncu launch list of a bench run and a --set full capture of one
# 256-frame pass of every extractor kernel.  usage: gpurun --timeout 1500 -- bash tools/round_final.sh TAG
tag=${1:-r1q}
python -m pytest tests -m gpu -x -q 2>&1 | tail -2
python bench.py --steps 20 --warmup 3 > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err
python bench.py --steps 20 --warmup 3 --chunk 1024 > gpurun_out/bench_${tag}_c1024.json 2> gpurun_out/bench_${tag}_c1024.err
for f in $tag ${tag}_c1024; do python -c "
import json
d=json.loads(open('gpurun_out/bench_$f.json').read().strip().splitlines()[-1]); print('$f', round(d['value']), round(d['ms_per_step'],3), round(d['e2e']['value']), d['gpu_launches'], d['roofline']['stage_ms_per_step'], d['clocks'])"; done
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/${tag}_launches.csv \
    python bench.py --frames 512 --unique 128 --chunk 256 --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_${tag}_l.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_level0|k_resize|k_border|k_fast|k_octree|k_blur|k_describe" -s 39 -c 13 \
    -f -o gpurun_out/${tag}_all_kernels python bench.py --frames 256 --unique 64 --chunk 256 --steps 2 --warmup 3 --no-cpu --no-match > gpurun_out/ncu_${tag}.log 2>&1
ls -la gpurun_out/${tag}_all_kernels.ncu-rep gpurun_out/${tag}_launches.csv
