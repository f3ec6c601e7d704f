python -m pytest tests/test_gpu_crops.py tests/test_gpu_nv12.py tests/test_gpu_facades.py -q -m gpu -x --no-header -p no:cacheprovider 2>&1 | tail -4
python scripts/k1_bench.py 2>&1 | grep K5
