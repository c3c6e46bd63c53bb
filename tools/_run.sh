timeout 300 python -m pytest tests/test_kernels_gpu.py -x -q -k "convnext_mlp" 2>&1 | tail -3
timeout 300 python tools/bench_mlp.py 2>&1 | tee gpurun_out/r2_mlp_fused_microbench.txt | tail -3
python bench.py --steps 20 --warmup 5 --no-extra 2>&1 | tail -1 > gpurun_out/r2_bench_tmp.json; python -c "
import json; b=json.loads(open('gpurun_out/r2_bench_tmp.json').read()); print(b['value'], b['e2e']['value'], b['sequential']['value'], b['roofline_dwconv'], b['roofline_mlp'])"
