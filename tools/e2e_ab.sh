# A/B of the host-buffer (e2e) path's pipeline parameters; usage: gpurun -- bash tools/e2e_ab.sh
run() { name=$1; shift; env "$@" python bench.py --steps 20 --warmup 3 $EXTRA > gpurun_out/ab_$name.json 2> gpurun_out/ab_$name.err; python -c "
import json
d=json.loads(open('gpurun_out/ab_$name.json').read().strip().splitlines()[-1]); print('$name', round(d['value']), round(d['e2e']['value']), round(d['e2e']['ms_per_step'],2), round(d['e2e']['copy_only_ms_per_step'],2))"; }
python -m pytest tests/test_gpu_extract.py -x -q 2>&1 | tail -1
EXTRA="--e2e-chunk 128" run s1c128 ORBX_E2E_STREAMS=1
EXTRA="--e2e-chunk 128" run s2c128 ORBX_E2E_STREAMS=2
EXTRA="--e2e-chunk 192" run s2c192 ORBX_E2E_STREAMS=2
EXTRA="--e2e-chunk 192" run s1c192 ORBX_E2E_STREAMS=1
