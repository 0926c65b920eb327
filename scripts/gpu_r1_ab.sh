( timeout 900 python -m pytest tests/test_fused_gpu.py tests/test_model_gpu.py tests/test_model_golden_gpu.py -x -q -m gpu ) 2>&1 | tail -2
python scripts/fused_micro.py e4m3 2>&1 | tail -7
for spec in e4m3 posit8_1; do
  python scripts/llama_bench.py --spec $spec --layers 8 --steps 10 --graph 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['spec'], 'ms', round(d['ms_per_window'],4), 'loss', d['loss'])"
done
python scripts/bert_bench.py 2>/dev/null | tail -1 | cut -c1-200
