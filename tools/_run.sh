python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py 2>gpurun_out/bench_n1.err | tail -1 > gpurun_out/r2_bench_n1.json
python -c "
import json; b=json.loads(open('gpurun_out/r2_bench_n1.json').read()); print(b['value'], b['e2e']['value'], b['sequential']['value'], b['roofline']['frac'], b['roofline']['traffic'], b['roofline_dwconv']['traffic'], b['roofline_mlp']['traffic'], b['clocks'])"
