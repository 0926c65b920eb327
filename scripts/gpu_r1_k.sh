# Round-end rehearsal: full GPU suite, smoke, reference arm, bench N=1 (what the driver runs).
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -x -q -m gpu ) > gpurun_out/pytest_gpu_full.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_full.log
tail -4 gpurun_out/pytest_gpu_full.log
( time python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log; tail -5 gpurun_out/smoke.log
( time python bench.py --impl reference ) > gpurun_out/bench_ref_k.json 2> gpurun_out/bench_ref_k.err; tail -c 600 gpurun_out/bench_ref_k.json
( time python bench.py ) > gpurun_out/bench_k.json 2> gpurun_out/bench_k.err; tail -5 gpurun_out/bench_k.err; cut -c1-400 gpurun_out/bench_k.json
