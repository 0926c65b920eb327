mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -x -q -m gpu ) > gpurun_out/pytest_gpu_full.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_full.log
tail -12 gpurun_out/pytest_gpu_full.log | cut -c1-600
( time python bench.py --no-llama --no-cpu-baseline --steps 10 ) > gpurun_out/bench_o.json 2> gpurun_out/bench_o.err; tail -3 gpurun_out/bench_o.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_o.json").read().strip().splitlines()[-1])
print(d["value"], d["roofline"]["frac"])
for k, v in d["other_shapes_GBps"].items():
    print(f"{v:8.1f}  {k}")
PY
