# Round-1 evidence run (one B200): launch list of the bench command, ncu --set full of the dominant kernels,
# per-kernel device times inside the captured Llama forward, micro-benchmarks.
set -x
mkdir -p gpurun_out
# (a) every launch of the bench command with its device time
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r01_bench.csv \
    python bench.py --steps 2 --warmup 1 --no-llama --no-cpu-baseline --no-extras > gpurun_out/bench_under_ncu.log 2>&1
# (b) the dominant kernel of the bench line, full set, at the bench's launch size (2^30 elements)
timeout 900 ncu --set full --clock-control none --import-source on -k regex:fq_flat -s 30 -c 8 -o gpurun_out/prof_r01_fq_flat \
    python bench.py --steps 1 --warmup 3 --no-llama --no-cpu-baseline --no-extras > gpurun_out/ncu_fq.log 2>&1
# (c) the GEMM kernel on Llama shapes (bf16 and fp8 operands)
timeout 900 ncu --set full --clock-control none --import-source on -k regex:qt_gemm -s 40 -c 6 -o gpurun_out/prof_r01_gemm \
    python scripts/gemm_bench.py > gpurun_out/ncu_gemm.log 2>&1
# (d) un-profiled numbers
python scripts/gemm_bench.py --json gpurun_out/gemm_bench_r01_final.json > gpurun_out/gemm_bench_r01_final.log 2>&1
python scripts/fused_micro.py e4m3 > gpurun_out/fused_micro_r01.log 2>&1; python scripts/fused_micro.py posit8_1 >> gpurun_out/fused_micro_r01.log 2>&1
python scripts/attn_micro.py > gpurun_out/attn_micro_r01.log 2>&1
python scripts/llama_profile.py --spec e4m3 > gpurun_out/llama_kernels_r01_e4m3.log 2>&1
python scripts/llama_profile.py --spec posit8_1 > gpurun_out/llama_kernels_r01_posit8_1.log 2>&1
python scripts/bert_bench.py > gpurun_out/bert_r01.json 2>/dev/null; python scripts/bert_bench.py --spec e4m3 >> gpurun_out/bert_r01.json 2>/dev/null
python scripts/bert_bench.py --no-fused >> gpurun_out/bert_r01.json 2>/dev/null
python scripts/bert_bench.py --model mobilebert-tiny --spec e4m3 --ops gemm,residual,layernorm,activation,scaling >> gpurun_out/bert_r01.json 2>/dev/null
python scripts/finetune_step.py > gpurun_out/finetune_r01_n1.json 2>/dev/null
tail -3 gpurun_out/gemm_bench_r01_final.log | cut -c1-200
