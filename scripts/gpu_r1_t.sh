mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests/test_fused_gpu.py tests/test_gemm_gpu.py tests/test_model_gpu.py tests/test_model_golden_gpu.py -x -q -m gpu ) > gpurun_out/pytest_causal.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_causal.log
tail -15 gpurun_out/pytest_causal.log | cut -c1-400
python scripts/llama_bench.py --spec e4m3 --layers 8 --steps 5 --graph 2>&1 | tail -2 | cut -c1-600
QT_CAUSAL=0 python scripts/llama_bench.py --spec e4m3 --layers 8 --steps 5 --graph 2>&1 | tail -1 | cut -c1-600
python scripts/llama_bench.py --spec posit8_1 --layers 8 --steps 5 --graph 2>&1 | tail -1 | cut -c1-600
QT_CAUSAL=0 python scripts/llama_bench.py --spec posit8_1 --layers 8 --steps 5 --graph 2>&1 | tail -1 | cut -c1-600
