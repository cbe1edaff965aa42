# usage: bash tools/_multi_n.sh N   -- stage times + bench line of the row-sharded step on N GPUs
N=$1
FOCAL_B200_STAGE_TIMES=1 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 tools/shard_stage_times.py 2>&1 | grep -E "rank|row-sharded" | tee gpurun_out/r2_shard_stage_times_$N.txt
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 100 --warmup 5 > gpurun_out/r2_bench_${N}gpu.json 2> gpurun_out/r2_bench_${N}gpu.err
python -c "
import json
d=json.loads(open('gpurun_out/r2_bench_${N}gpu.json').read().strip().splitlines()[-1])
print({k:d.get(k) for k in ('n_gpus','value','ms_per_step','loss')}); print(d['e2e'])"
