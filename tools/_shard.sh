for R in 1 2 4 8; do python tools/shard_probe.py 0 $R 50 2>&1 | tail -1; done
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r2_shard8_launches.csv python tools/shard_probe.py 3 8 3 > /dev/null 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/r2_shard8_launches.csv')) if len(r)>5]
hdr=rows[0]; ik=hdr.index('Kernel Name'); iv=hdr.index('Metric Value')
last={}
order=[]
for r in rows[1:]:
    k=r[ik].split('(')[0][:60]
    if k not in last: order.append(k)
    last[k]=float(r[iv].replace(',',''))
for k in order: print(f"{last[k]/1000:8.1f} us  {k}")
PY
