timeout 600 python -m pytest tests/test_conv_gemm_gpu.py -x -q -k whole_range 2>&1 | grep -E "Error|assert|tensor|passed|failed" | head -12
