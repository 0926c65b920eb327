mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests/test_fq_gpu.py tests/test_block_gpu.py -x -q -m gpu ) > gpurun_out/pytest_fq_block.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_fq_block.log
tail -6 gpurun_out/pytest_fq_block.log | cut -c1-600
( time python bench.py --no-llama --no-cpu-baseline --steps 10 ) > gpurun_out/bench_p.json 2> gpurun_out/bench_p.err; tail -3 gpurun_out/bench_p.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_p.json").read().strip().splitlines()[-1])
print(d["value"], d["roofline"]["frac"])
for k, v in d["other_shapes_GBps"].items():
    print(f"{v:8.1f}  {k}")
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"mx_flat|gwa_flat|mx_cols" -c 6 -o gpurun_out/prof_r01_mx_v2 python scripts/mx_micro.py --reps 1 --only "bs=" > gpurun_out/ncu_mx2.log 2>&1; tail -2 gpurun_out/ncu_mx2.log
