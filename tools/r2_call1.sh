#!/bin/bash
# Round 2, GPU call 1: (a) parity + A/B of the four-pair FAST quick test, (b) first numbers of BASELINE configs C1 / C3 / C5,
# (c) ncu launch list + --set full capture of one C5 (3840x2160 / 8000 / 12) device pass.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
ORBX_FAST_PAIR4=1 python -m pytest tests/test_gpu_extract.py tests/test_gpu_fullsize.py -x -q 2>&1 | tail -3 | tee gpurun_out/r2a_pair4_tests.txt
for v in 0 1 0 1; do ORBX_FAST_PAIR4=$v python bench.py --steps 10 --warmup 3 --no-match --no-cpu --no-configs 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('pair4=$v', round(d['value']), d['roofline']['stage_ms_per_step'])" | tee -a gpurun_out/r2a_pair4_ab.txt; done
for c in C1 C3 C5; do
  python bench.py --only-config $c --steps 5 --no-cpu > gpurun_out/r2a_cfg_$c.json 2> gpurun_out/r2a_cfg_$c.err || tail -5 gpurun_out/r2a_cfg_$c.err
  cat gpurun_out/r2a_cfg_$c.json
done
# C5 pass under ncu: launch list first, then the full set on one pass (17 launches) after the warm-up passes
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2a_c5_launches.csv \
    python bench.py --only-config C5 --c5-frames 8 --c5-batch 8 --steps 1 --no-cpu > gpurun_out/r2a_c5_ncu_list.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -s 51 -c 17 -f -o gpurun_out/r2a_c5_pass \
    python bench.py --only-config C5 --c5-frames 8 --c5-batch 8 --steps 1 --no-cpu > gpurun_out/r2a_c5_ncu_full.log 2>&1
ls -la gpurun_out/*.ncu-rep
