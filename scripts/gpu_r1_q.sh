mkdir -p gpurun_out /tmp/prof
prof() {  # name, kernel regex, command...
  name=$1; shift; k=$1; shift
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:$k -c 1 -o /tmp/prof/$name "$@" > /tmp/prof/$name.log 2>&1
  python scripts/ncu_summarize.py /tmp/prof/$name.ncu-rep > gpurun_out/ncu_r01_$name.txt 2>&1
  head -8 gpurun_out/ncu_r01_$name.txt | cut -c1-150
}
REPS=1 prof fq_cols fq_cols python scripts/fq_cols_one.py
prof mx_flat_fp4 mx_flat python scripts/mx_micro.py --reps 1 --only "fp4_e2m1,qs=microscaling,bs=32,ax=-1"
prof mx_flat_int6 mx_flat python scripts/mx_micro.py --reps 1 --only "int6,qs=microscaling,bs=64,ax=-1"
prof gwa_flat gwa_flat python scripts/mx_micro.py --reps 1 --only "uint4,qs=group_wise_affine"
prof mx_cols mx_cols python scripts/mx_micro.py --reps 1 --only "int6,qs=microscaling,bs=64,ax=0"
ls -la gpurun_out
