mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_block_gpu.py -x -q -m gpu ) > gpurun_out/pytest_block.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_block.log
tail -12 gpurun_out/pytest_block.log | cut -c1-600
python scripts/mx_micro.py > gpurun_out/mx_micro_b.log 2>&1; cat gpurun_out/mx_micro_b.log
