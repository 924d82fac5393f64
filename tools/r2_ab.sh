#!/bin/bash
# A/B of an environment switch on the short bench line.  usage: r2_ab.sh VAR v1 v2 ...
var=$1; shift
for v in "$@" "$@"; do env $var=$v python bench.py --steps 10 --warmup 3 --no-match --no-cpu --no-configs 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$var=$v', round(d['value']), {k: round(x,3) for k,x in d['roofline']['stage_ms_per_step'].items()})"; done
