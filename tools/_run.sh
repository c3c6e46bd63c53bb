for w in 4 8; do echo "UC_DW_MMA_WARPS=$w"; UC_DW_MMA_WARPS=$w timeout 300 python tools/bench_dw.py 2>&1 | sed -e 's/static-schedule.*//' ; done
UC_DW_MMA_WARPS=8 timeout 300 python -m pytest tests/test_kernels_gpu.py -x -q -k "dwconv_tiled" 2>&1 | tail -2
timeout 300 python -m pytest tests/test_kernels_gpu.py -x -q -k "dwconv_tiled" 2>&1 | tail -2
