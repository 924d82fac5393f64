#!/bin/bash
# Round 2: parity of the extractor GPU tests + a short bench line (+ optional config blocks).  usage: r2_check.sh <tag> [configs...]
tag=${1:-r2x}; shift
mkdir -p gpurun_out
python -m pytest tests/test_gpu_extract.py tests/test_gpu_fullsize.py -x -q 2>&1 | tail -3 | tee gpurun_out/tests_$tag.txt
grep -q "failed\|error" gpurun_out/tests_$tag.txt && exit 1
python bench.py --steps 10 --warmup 3 --no-match --no-cpu --no-configs > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err
python -c "
import json
d=json.loads(open('gpurun_out/bench_$tag.json').read().strip().splitlines()[-1]); print('$tag', round(d['value']), round(d['ms_per_step'],3), round(d['e2e']['value']), d['roofline']['stage_ms_per_step'])"
for c in "$@"; do
  python bench.py --only-config $c --steps 5 --no-cpu > gpurun_out/cfg_${tag}_$c.json 2> gpurun_out/cfg_${tag}_$c.err || tail -5 gpurun_out/cfg_${tag}_$c.err
  python -c "
import json
d=json.loads(open('gpurun_out/cfg_${tag}_$c.json').read().strip().splitlines()[-1])['$c']; print('$c', {k: d[k] for k in ('frames_per_s','stage_ms_per_frame') if k in d} or d)"
done
