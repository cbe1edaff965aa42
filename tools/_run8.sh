cd /root/repo
N=$1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
for wl in headline cfg4 cfg5; do
  timeout 400 $TR bench.py --gpus $N --steps 30 --warmup 5 --workload $wl > gpurun_out/bench_v6_${wl}_$N.json 2> gpurun_out/bench_v6_${wl}_$N.err
  tail -c 300 gpurun_out/bench_v6_${wl}_$N.err | grep -v OMP
  python -c "
import json;j=json.loads(open('gpurun_out/bench_v6_${wl}_$N.json').read().strip().splitlines()[-1]);print('$wl N=$N',j['value'],j['ms_per_step'],j['tensor_roofline_frac'],j['e2e']['value'],j['gpu_launches'],j['config']['parallelism'][:40])"
done
