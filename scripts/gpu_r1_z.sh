( timeout 900 python -m pytest tests/test_fq_gpu.py tests/test_block_gpu.py tests/test_model_golden_gpu.py -x -q -m gpu ) 2>&1 | tail -3
python bench.py --no-llama --no-cpu-baseline --steps 10 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['value'],1), d['roofline']['frac'], 'e2e', round(d['e2e']['value'],1))
for k,v in d['other_shapes_GBps'].items():
    if 'scale' in k: print(round(v), k)"
