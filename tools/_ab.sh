# usage: bash tools/_ab.sh ENVVAR  -- headline / cfg2 bench lines with ENVVAR=1 and =0 on the same box, twice each
V=$1
for rep in 1 2; do for wl in headline cfg2; do for x in 1 0; do
env $V=$x python bench.py --workload $wl --steps 100 --warmup 5 --no-cpu-baseline --sustain-seconds 0 > gpurun_out/_ab.json 2> gpurun_out/_ab.err
python - <<PY
import json
d=json.loads(open('gpurun_out/_ab.json').read().strip().splitlines()[-1])
print('$wl $V=$x', round(d['ms_per_step']*1e3,1), 'us', d['loss'])
PY
done; done; done | tee gpurun_out/r2_ab_$V.txt
