mkdir -p gpurun_out /tmp/prof
timeout 300 ncu --set full --clock-control none --import-source on -k regex:softmax_fq -s 30 -c 1 -o /tmp/prof/softmax python scripts/causal_micro.py > /tmp/prof/softmax.log 2>&1
python scripts/ncu_summarize.py /tmp/prof/softmax.ncu-rep > gpurun_out/ncu_r01_softmax_e4m3.txt 2>&1
head -30 gpurun_out/ncu_r01_softmax_e4m3.txt | cut -c1-160
