#!/bin/bash
# Round-2 starting point: validates and times the experimental four-pair quick test of k_fast_tma (ORBX_FAST_PAIR4=1, not yet run on
# a GPU at the end of round 1).  usage: gpurun --timeout 300 -- bash tools/pair4_check.sh
export ORBX_FAST_PAIR4=1
python -m pytest tests/test_gpu_extract.py tests/test_gpu_fullsize.py -x -q 2>&1 | tail -2
for v in 0 1; do ORBX_FAST_PAIR4=$v python bench.py --steps 10 --warmup 3 --no-match --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('pair4=$v', round(d['value']), d['roofline']['stage_ms_per_step'])"; done
