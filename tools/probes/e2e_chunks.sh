for c in 96 128 192 256; do python bench.py --no-cpu --no-match --steps 8 --e2e-chunk $c 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('e2e_chunk=$c', round(d['value']), round(d['e2e']['value']), round(d['e2e']['ms_per_step'],2))"; done
