#!/bin/bash
# describe warps per SM A/B (ORBX_DESC_WARPS caps the warps of the one CTA per SM)
python -m pytest tests/test_gpu_extract.py tests/test_gpu_fullsize.py tests/test_gpu_stereo.py -x -q 2>&1 | tail -3
line() { python bench.py --steps 10 --warmup 3 --no-match --no-cpu --no-configs 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1', round(d['value']), round(d['e2e']['value']), {k: round(v,3) for k,v in d['roofline']['stage_ms_per_step'].items()})"; }
for w in 24 0 26 0; do ORBX_DESC_WARPS=$w line "desc_warps=$w" | tee -a gpurun_out/dw_ab.txt; done
python bench.py --only-config C1 --steps 5 --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])['C1']; print('C1', d['pageable_caller_buffers']['median_ms'], d['pinned_caller_buffers']['median_ms'])" | tee -a gpurun_out/dw_ab.txt
python bench.py --only-config C5 --steps 5 --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])['C5']; print('C5', round(d['frames_per_s']), d['stage_ms_per_frame'])" | tee -a gpurun_out/dw_ab.txt
