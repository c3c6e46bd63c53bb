timeout 900 python -m pytest tests/test_vos_gpu.py -x -q 2>&1 | tail -8
python bench.py --steps 24 --warmup 5 2>gpurun_out/bench.err | tail -1 > gpurun_out/r2_bench_tmp.json; tail -3 gpurun_out/bench.err; python -c "
import json; b=json.loads(open('gpurun_out/r2_bench_tmp.json').read()); print(b['value'], b['e2e']['value']); print({k:(round(b[k]['value'],1), round(b[k]['e2e'],1)) for k in ['mot_1536x2048','vos_800x1280_1obj','vos_800x1280_3obj']})"
