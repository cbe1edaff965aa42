set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee gpurun_out/r2_gputest_full.txt
python bench.py --steps 50 --warmup 5 --no-cpu-baseline --sustain-seconds 0 > gpurun_out/r2_bench_e.json 2> gpurun_out/r2_bench_e.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench_e.json').read().strip().splitlines()[-1])
print(d['ms_per_step'], d['stages_ms'], d['loss'], d['e2e'])
PY
