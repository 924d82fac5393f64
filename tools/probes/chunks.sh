for c in 128 256 512 1024; do python bench.py --no-cpu --no-match --steps 6 --chunk $c 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('chunk=$c', round(d['value']), round(d['e2e']['value']), {k: round(x,2) for k,x in d['roofline']['stage_ms_per_step'].items()})"; done
