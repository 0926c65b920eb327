set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -25 gpurun_out/pytest_gpu.log | cut -c1-600
timeout 600 python scripts/llama_bench.py --spec e4m3 --steps 5 --graph > gpurun_out/llama_e4m3_graph.json 2> gpurun_out/llama.err; tail -5 gpurun_out/llama.err; cat gpurun_out/llama_e4m3_graph.json
