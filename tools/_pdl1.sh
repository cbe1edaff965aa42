# 1 GPU: full GPU test suite, then bench lines with / without programmatic dependent launch
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/r2_gputest_full.txt
for wl in headline cfg2 cfg1; do for pdl in 1 0; do
FOCAL_B200_PDL=$pdl python bench.py --workload $wl --steps 100 --warmup 5 --no-cpu-baseline --sustain-seconds 0 > gpurun_out/r2_pdl_${wl}_$pdl.json 2> gpurun_out/r2_pdl_${wl}_$pdl.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r2_pdl_${wl}_$pdl.json').read().strip().splitlines()[-1])
print('$wl pdl=$pdl', round(d['ms_per_step']*1e3,1), 'us', {k:round(v*1e3,1) for k,v in (d['stages_ms'] or {}).items()}, d['loss'], round(d['e2e']['ms_per_step']*1e3,1))
PY
done; done | tee gpurun_out/r2_pdl_ab.txt
