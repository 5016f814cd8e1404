timeout 300 python -m pytest tests/test_gpu_train.py -q -m gpu -x --timeout=120 2>&1 | tail -15
timeout 200 python tools/train_bench.py 2>&1 | head -12
