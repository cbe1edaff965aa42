# last check of the tree as committed: GPU tests, smoke, the driver's bench invocation (own arm + reference arm)
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py --gpus 1 --steps 20 --warmup 3 > gpurun_out/r2_last_bench.json 2> gpurun_out/r2_last_bench.err
python -c "
import json
d=json.loads(open('gpurun_out/r2_last_bench.json').read().strip().splitlines()[-1])
print({k:d.get(k) for k in ('value','ms_per_step','gpu_launches','steps','warmup','loss')}); print(d['e2e']); print(d['roofline']['frac'], d['roofline']['executed_frac'], d['cpu_baseline']['value'])"
