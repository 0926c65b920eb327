python bench.py --no-llama --no-cpu-baseline --steps 10 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['value'],1), d['roofline']['frac']); print({k:round(v) for k,v in d['roofline']['per_spec_GBps'].items()})
for k,v in d['other_shapes_GBps'].items():
    if 'direct' in k or 'int8' in k or 'microscaling' in k: print(round(v), k)"
