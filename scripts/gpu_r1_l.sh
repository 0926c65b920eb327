mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_block_gpu.py -x -q -m gpu ) > gpurun_out/pytest_block.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_block.log
tail -30 gpurun_out/pytest_block.log | cut -c1-400
( time python bench.py --no-llama --no-cpu-baseline --steps 10 ) > gpurun_out/bench_l.json 2> gpurun_out/bench_l.err; tail -3 gpurun_out/bench_l.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_l.json").read().strip().splitlines()[-1])
for k, v in d["other_shapes_GBps"].items():
    print(f"{v:8.1f}  {k}")
PY
