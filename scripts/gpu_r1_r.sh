mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests/test_fq_gpu.py tests/test_block_gpu.py -x -q -m gpu ) > gpurun_out/pytest_fq_block.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_fq_block.log
tail -4 gpurun_out/pytest_fq_block.log | cut -c1-600
python scripts/mx_micro.py > gpurun_out/mx_micro_c.log 2>&1; cat gpurun_out/mx_micro_c.log
python scripts/fq_cols_one.py
python scripts/fq_cols_one.py e4m3
