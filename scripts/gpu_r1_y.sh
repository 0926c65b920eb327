mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -x -q -m gpu ) > gpurun_out/pytest_gpu_pdl.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_pdl.log
tail -5 gpurun_out/pytest_gpu_pdl.log | cut -c1-300
for pdl in 1 0; do
  for spec in e4m3 posit8_1; do
    QT_PDL=$pdl python scripts/llama_bench.py --spec $spec --layers 8 --steps 10 --graph 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('PDL=$pdl', d['spec'], 'ms', round(d['ms_per_window'],4), 'loss', d['loss'])"
  done
done
QT_PDL=1 python scripts/bert_bench.py 2>/dev/null | tail -1 | cut -c1-200
QT_PDL=0 python scripts/bert_bench.py 2>/dev/null | tail -1 | cut -c1-200
