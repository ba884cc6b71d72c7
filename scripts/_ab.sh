python -m pytest tests -m gpu -q -x 2>&1 | tail -3
run() { python bench.py --no-transmil --no-cpu-baseline --no-cls-row-only 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    l=l.strip()
    if l.startswith('{'):
        d=json.loads(l); print('$1', round(d['ms_per_step'],4), round(d['e2e']['ms_per_step'],4), round(d['sustained']['ms_per_step'],4), d['gpu_launches'])
"; }
run plan; DML_B200_FWD_HALF_BLOCKS=0 run nosplit; run plan; DML_B200_FWD_HALF_BLOCKS=0 run nosplit
