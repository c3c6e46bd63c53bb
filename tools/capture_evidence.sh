#!/bin/bash
# Round-2 evidence for profiles/ (run on a B200 through gpurun; everything lands in gpurun_out/ and is then summarised by
# tools/launch_summary.py, tools/conv_table.py, tools/ncu_summary.py, tools/make_roofline_inputs.py — see profiles/README.md).
# Numbers printed by a run under ncu are never bench values.
O=gpurun_out
NCU="ncu --clock-control none"
# 1. every kernel of init + 2 steady-state 800x1280 SOT frames with its device time (caches warm between kernels, kernels serialised by ncu)
UC_CONV_TRACE=$O/r2_conv_trace.json $NCU --metrics gpu__time_duration.sum --cache-control none --csv --log-file $O/r2_launches_warm_final.csv \
  python tools/profile_frame.py unicorn_track_large 2 > $O/profile_frame.log 2>&1
# 2. DRAM bytes of every kernel of the same run
$NCU --metrics dram__bytes_read.sum,dram__bytes_write.sum --cache-control none --csv --log-file $O/r2_frame_dram_bytes.csv \
  python tools/profile_frame.py unicorn_track_large 2 > /dev/null 2>&1
# 3. ncu --set full of the hand-written hot kernels (one launch each, cold L2)
$NCU --set full --import-source on -k regex:dwconv7_mma -s 5 -c 1 -f -o $O/r2_ncu_dwmma_s1 python tools/bench_dw.py s1 > /dev/null 2>&1
$NCU --set full --import-source on -k regex:dwconv7_mma -s 5 -c 1 -f -o $O/r2_ncu_dwmma_s3 python tools/bench_dw.py s3 > /dev/null 2>&1
$NCU --set full --import-source on -k regex:convnext_mlp -s 3 -c 1 -f -o $O/r2_ncu_mlp_s1 python tools/bench_mlp.py > /dev/null 2>&1
ls -la $O/*.ncu-rep
