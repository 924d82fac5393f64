#!/bin/bash
# Round 2 state check: the whole gpu test-suite, the default bench line exactly as the driver runs it (timed), the reference arm.
# usage: gpurun --timeout 1500 -- bash tools/r2_state.sh TAG
tag=${1:-r2s}
mkdir -p gpurun_out
[ -n "$SKIP_TESTS" ] || python -m pytest tests -m gpu -q 2>&1 | tail -4 | tee gpurun_out/tests_$tag.txt
t0=$(date +%s.%N); python bench.py > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err; echo "bench wall $(python -c "import time; print(round(time.time() - $t0, 1))") s"
t0=$(date +%s.%N); python bench.py --impl reference > gpurun_out/bench_${tag}_ref.json 2> gpurun_out/bench_${tag}_ref.err; echo "reference arm wall $(python -c "import time; print(round(time.time() - $t0, 1))") s"
python -c "
import json
d=json.loads(open('gpurun_out/bench_$tag.json').read().strip().splitlines()[-1]); r=json.loads(open('gpurun_out/bench_${tag}_ref.json').read().strip().splitlines()[-1])
print('$tag', round(d['value']), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value']), 'ref', round(r['value'],1), d['gpu_launches'], d['roofline']['stage_ms_per_step'], d['clocks'])"
