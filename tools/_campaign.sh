# round-2 measurement campaign on one B200 (final build): bench lines of every BASELINE.json config, reference arm,
# ncu launch list of the bench command, ncu --set full of the Gram and row kernels
set -x
O=gpurun_out
python bench.py > $O/r2_final_headline.json 2> $O/r2_final_headline.err
python bench.py --impl reference --steps 3 --warmup 1 > $O/r2_final_reference.json 2> $O/r2_final_reference.err
python bench.py --precision fp32 --no-cpu-baseline > $O/r2_final_headline_fp32.json 2> $O/r2_final_headline_fp32.err
for w in cfg1 cfg2 cfg3 cfg4 cfg5 cfg5w; do
  python bench.py --workload $w --no-cpu-baseline --sustain-seconds 0 > $O/r2_final_$w.json 2> $O/r2_final_$w.err
done
python bench.py --workload cfg2 --precision fp32 --no-cpu-baseline --sustain-seconds 0 > $O/r2_final_cfg2_fp32.json 2> $O/r2_final_cfg2_fp32.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 150 --csv --log-file $O/r2_launches_final.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --sustain-seconds 0 > $O/b_ncu.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:"gram_kernel|_v3_kernel" -s 10 -c 5 -o $O/r2_prof_final -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --sustain-seconds 0 > $O/b_ncu2.log 2>&1
tail -2 $O/b_ncu2.log
for f in headline headline_fp32 cfg1 cfg2 cfg3 cfg4 cfg5 cfg5w cfg2_fp32 reference; do python - $O/r2_final_$f.json <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1].split('/')[-1], d.get('ms_per_step'), d.get('value'), (d.get('roofline') or {}).get('frac'), d.get('tensor_roofline_frac'), (d.get('e2e') or {}).get('value'))
except Exception as e: print(sys.argv[1], 'ERR', e)
PY
done
