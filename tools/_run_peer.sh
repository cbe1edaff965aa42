cd /root/repo
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port 29511"
timeout 300 $TR tools/dist_gpu_check.py 2>&1 | grep -v "^W\|^\*\*\*" | tail -30
for peer in 1 0; do
  FOCAL_B200_PEER=$peer timeout 300 $TR bench.py --gpus $1 --steps 50 --warmup 5 > gpurun_out/bench_peer${peer}_$1.json 2> gpurun_out/bench_peer${peer}_$1.err
  tail -c 400 gpurun_out/bench_peer${peer}_$1.err
  python -c "
import json;j=json.loads(open('gpurun_out/bench_peer${peer}_$1.json').read().strip().splitlines()[-1]);print('peer=$peer',j['value'],j['ms_per_step'],j.get('host_enqueue_ms_per_step'),j['e2e'],j['gpu_launches'])"
done
