bash tools/capture_evidence.sh > gpurun_out/capture.log 2>&1
timeout 2400 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
python bench.py 2>gpurun_out/bench_n1.err | tail -1 > gpurun_out/r2_bench_n1.json
python bench.py --impl reference --steps 3 --warmup 1 2>/dev/null | tail -1 > gpurun_out/r2_bench_reference_arm.json
python -c "
import json; b=json.loads(open('gpurun_out/r2_bench_n1.json').read()); print(b['value'], b['e2e']['value'], b['sequential']['value'], b['roofline']['frac'], b['cpu_baseline'])"
cat gpurun_out/r2_bench_reference_arm.json | cut -c1-400
