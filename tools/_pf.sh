timeout 300 python -m pytest tests/test_gpu_lookup_conv.py -x -q -s 2>&1 | tail -40
