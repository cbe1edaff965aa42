cd /root/repo
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port 29511"
timeout 300 $TR tools/dist_gpu_check.py 2>&1 | grep "rank 0/" | tail -8
timeout 300 $TR bench.py --gpus $1 --steps 50 --warmup 5 > gpurun_out/bench_v7_$1.json 2> gpurun_out/bench_v7_$1.err
tail -c 300 gpurun_out/bench_v7_$1.err | grep -v "OMP\|\*\*\*"
python -c "
import json;j=json.loads(open('gpurun_out/bench_v7_$1.json').read().strip().splitlines()[-1]);print('N=$1',j['value'],j['ms_per_step'],j.get('host_enqueue_ms_per_step'),j['e2e']['value'],j['gpu_launches'])"
