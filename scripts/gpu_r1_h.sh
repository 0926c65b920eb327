set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gemm_gpu.py tests/test_model_gpu.py -m gpu -q -x > gpurun_out/pytest_gemm.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gemm.log
tail -15 gpurun_out/pytest_gemm.log | cut -c1-400
timeout 600 python scripts/gemm_bench.py --json gpurun_out/gemm_bench.json > gpurun_out/gemm_bench.log 2>&1; cat gpurun_out/gemm_bench.log | grep -v "^{" | cut -c1-200
timeout 600 python scripts/llama_bench.py --spec posit8_1 --steps 5 --graph > gpurun_out/llama_posit_graph.json 2> gpurun_out/llama.err; tail -3 gpurun_out/llama.err; cat gpurun_out/llama_posit_graph.json
timeout 600 python scripts/llama_bench.py --spec e4m3 --steps 5 --graph > gpurun_out/llama_e4m3_graph.json 2>> gpurun_out/llama.err; tail -3 gpurun_out/llama.err; cat gpurun_out/llama_e4m3_graph.json
