set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -25 gpurun_out/pytest_gpu.log | cut -c1-400
timeout 600 python scripts/llama_bench.py --spec posit8_1 --steps 5 > gpurun_out/llama_posit_eager.json 2> gpurun_out/llama.err; tail -3 gpurun_out/llama.err; cat gpurun_out/llama_posit_eager.json
timeout 600 python scripts/llama_bench.py --spec posit8_1 --steps 5 --torch-gemm > gpurun_out/llama_posit_eager_cublas.json 2>> gpurun_out/llama.err; cat gpurun_out/llama_posit_eager_cublas.json
timeout 600 python scripts/llama_bench.py --spec posit8_1 --steps 5 --graph > gpurun_out/llama_posit_graph.json 2>> gpurun_out/llama.err; tail -3 gpurun_out/llama.err; cat gpurun_out/llama_posit_graph.json
timeout 600 python scripts/llama_bench.py --spec e4m3 --steps 5 --graph > gpurun_out/llama_e4m3_graph.json 2>> gpurun_out/llama.err; cat gpurun_out/llama_e4m3_graph.json
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?" >> gpurun_out/bench.err
python -c "
import json; d=json.load(open('gpurun_out/bench.json')); print(d['value'], d['roofline']['frac'], d['e2e'], d['cpu_baseline']); print(d['other_shapes_GBps'])"
