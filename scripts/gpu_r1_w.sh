# Round-1 closing evidence (one B200): full GPU suite, smoke, both bench arms, launch list of the bench command.
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -x -q -m gpu ) > gpurun_out/pytest_gpu_full.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_full.log
tail -4 gpurun_out/pytest_gpu_full.log
( time python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log; tail -3 gpurun_out/smoke.log
( time python bench.py --impl reference ) > gpurun_out/bench_ref_w.json 2> gpurun_out/bench_ref_w.err
( time python bench.py ) > gpurun_out/bench_w.json 2> gpurun_out/bench_w.err; tail -3 gpurun_out/bench_w.err; cut -c1-300 gpurun_out/bench_w.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r01_bench_final.csv \
    python bench.py --steps 2 --warmup 1 --no-llama --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
wc -l gpurun_out/launches_r01_bench_final.csv
