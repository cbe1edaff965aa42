# 1 GPU: full GPU test suite + headline / cfg2 bench lines
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/r2_gputest_full.txt
for wl in headline cfg2; do
python bench.py --workload $wl --steps 100 --warmup 5 --no-cpu-baseline --sustain-seconds 0 > gpurun_out/r2_t1_${wl}.json 2> gpurun_out/r2_t1_${wl}.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r2_t1_${wl}.json').read().strip().splitlines()[-1])
print('$wl', round(d['ms_per_step']*1e3,1), 'us', {k:round(v*1e3,1) for k,v in (d['stages_ms'] or {}).items()}, d['loss'], round(d['e2e']['ms_per_step']*1e3,1))
PY
done | tee gpurun_out/r2_t1.txt
