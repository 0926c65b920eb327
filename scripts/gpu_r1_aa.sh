mkdir -p gpurun_out /tmp/prof
for k in rope_fq norm_fq fq_transpose; do
  QT_PDL=0 timeout 300 ncu --set full --clock-control none --import-source on -k regex:$k -s 3 -c 1 -o /tmp/prof/$k python scripts/fused_micro.py e4m3 > /tmp/prof/$k.log 2>&1
  python scripts/ncu_summarize.py /tmp/prof/$k.ncu-rep > gpurun_out/ncu_r01_$k.txt 2>&1
  head -24 gpurun_out/ncu_r01_$k.txt | grep -E "Kernel Name|duration|dram__bytes|inst_executed.sum|issue_active|warps_active|registers|grid_size|long_scoreboard|barrier" | cut -c1-140
done
