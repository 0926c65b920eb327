set -x
mkdir -p gpurun_out
timeout 600 python scripts/llama_bench.py --spec posit8_1 --steps 5 > gpurun_out/llama_posit_eager.json 2> gpurun_out/llama.err; tail -3 gpurun_out/llama.err; cat gpurun_out/llama_posit_eager.json
timeout 600 python scripts/llama_bench.py --spec posit8_1 --steps 5 --graph > gpurun_out/llama_posit_graph.json 2>> gpurun_out/llama.err; tail -5 gpurun_out/llama.err; cat gpurun_out/llama_posit_graph.json
timeout 600 python scripts/llama_bench.py --spec posit8_1 --steps 5 --graph --batch 8 > gpurun_out/llama_posit_graph_b8.json 2>> gpurun_out/llama.err; tail -3 gpurun_out/llama.err; cat gpurun_out/llama_posit_graph_b8.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_llama2l.csv python scripts/llama_bench.py --spec posit8_1 --steps 1 --layers 2 > gpurun_out/llama_ncu.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:qt_gemm -s 20 -c 6 -o gpurun_out/prof_gemm_r01 python scripts/gemm_bench.py > gpurun_out/ncu_gemm.log 2>&1
ls -la gpurun_out
