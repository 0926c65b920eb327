mkdir -p gpurun_out
python scripts/mx_micro.py > gpurun_out/mx_micro_a.log 2>&1; cat gpurun_out/mx_micro_a.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:mx_flat -c 3 -o gpurun_out/prof_r01_mx_flat python scripts/mx_micro.py --reps 1 --only "bs=32,ax=-1" > gpurun_out/ncu_mx.log 2>&1; tail -2 gpurun_out/ncu_mx.log
