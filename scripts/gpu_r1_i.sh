set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_model_gpu.py -m gpu -q -x > gpurun_out/pytest_model.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_model.log
tail -25 gpurun_out/pytest_model.log | cut -c1-300
for spec in posit8_1 e4m3; do
timeout 600 python scripts/llama_bench.py --spec $spec --steps 5 --graph > gpurun_out/llama_${spec}_fused.json 2> gpurun_out/llama.err; tail -3 gpurun_out/llama.err; cat gpurun_out/llama_${spec}_fused.json
done
timeout 600 python scripts/llama_bench.py --spec posit8_1 --steps 5 --graph --no-fused > gpurun_out/llama_posit8_1_unfused.json 2> gpurun_out/llama.err; tail -3 gpurun_out/llama.err; cat gpurun_out/llama_posit8_1_unfused.json
timeout 600 python scripts/llama_bench.py --spec posit8_1 --steps 5 > gpurun_out/llama_posit8_1_fused_eager.json 2> gpurun_out/llama.err; tail -3 gpurun_out/llama.err; cat gpurun_out/llama_posit8_1_fused_eager.json
