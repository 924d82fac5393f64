#!/bin/bash
# usage: tools/stage_sweep.sh ENVVAR v1 v2 ...   -> frames/s and stage times of bench.py for each value of the environment variable
var=$1; shift
for v in "$@"; do
  env $var=$v python bench.py --no-cpu --no-match --steps 6 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$var=$v', round(d['value']), {k: round(x,2) for k,x in d['roofline']['stage_ms_per_step'].items()})"
done
