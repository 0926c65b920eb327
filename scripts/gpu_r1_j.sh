mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_llama2l_fused.csv python scripts/llama_bench.py --spec posit8_1 --layers 2 --steps 1 > gpurun_out/llama_ncu.log 2>&1
tail -2 gpurun_out/llama_ncu.log | cut -c1-300
python scripts/debug_fused.py 2>&1 | tail -12
