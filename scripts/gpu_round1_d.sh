set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gemm_gpu.py -m gpu -q -x > gpurun_out/pytest_gemm.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gemm.log
tail -30 gpurun_out/pytest_gemm.log
nvidia-smi --query-gpu=name,memory.used --format=csv
timeout 600 python -m pytest tests/test_fq_gpu.py -m gpu -q > gpurun_out/pytest_fq.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_fq.log
tail -5 gpurun_out/pytest_fq.log
timeout 300 python scripts/gemm_bench.py > gpurun_out/gemm_bench.log 2>&1; cat gpurun_out/gemm_bench.log | head -20
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?" >> gpurun_out/bench.err
cat gpurun_out/bench.json | cut -c1-1500
