#!/bin/bash
# Validation of the last FAST / blur change: extractor gpu tests, a short bench line, and (only if the tests pass) a --set full
# capture of one k_fast_tma and one k_blur launch.
tag=${1:-r1t}
python -m pytest tests/test_gpu_extract.py tests/test_gpu_fullsize.py -x -q 2>&1 | tail -2 | tee gpurun_out/tests_$tag.txt
grep -q "failed\|error" gpurun_out/tests_$tag.txt && exit 1
python bench.py --steps 10 --warmup 3 --no-match --no-cpu > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err
python -c "
import json
d=json.loads(open('gpurun_out/bench_$tag.json').read().strip().splitlines()[-1]); print('$tag', round(d['value']), round(d['ms_per_step'],3), round(d['e2e']['value']), d['roofline']['stage_ms_per_step'])"
timeout 60 ncu --set full --clock-control none --import-source on -k regex:"k_fast_tma|k_blur" -s 6 -c 2 \
    -f -o gpurun_out/${tag}_fast_blur python bench.py --frames 256 --unique 64 --chunk 256 --steps 1 --warmup 3 --no-cpu --no-match > gpurun_out/ncu_${tag}.log 2>&1
ls -la gpurun_out/${tag}_fast_blur.ncu-rep
