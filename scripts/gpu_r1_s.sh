mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_fq_gpu.py -x -q -m gpu -k "host_streaming" ) 2>&1 | tail -3
for c in 20 22 24; do
  python bench.py --no-llama --no-cpu-baseline --no-extras --steps 5 --e2e-log2-chunk $c 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('chunk log2', $c, 'e2e', round(d['e2e']['value'],1), 'value', round(d['value'],1))"
done
python bench.py --no-llama --no-cpu-baseline --no-extras --steps 5 --e2e-log2-chunk 22 --e2e-log2-numel 28 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('numel 2^28 chunk 22 e2e', round(d['e2e']['value'],1))"
nvidia-smi --query-gpu=pcie.link.gen.current,pcie.link.width.current,pcie.link.gen.max --format=csv
