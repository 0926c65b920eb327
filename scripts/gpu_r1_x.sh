( timeout 900 python -m pytest tests/test_fused_gpu.py tests/test_model_gpu.py -x -q -m gpu ) 2>&1 | tail -3
python scripts/causal_micro.py 2>&1 | grep -E "QK|mask check"
