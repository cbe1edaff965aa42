# usage: bash tools/_multi2.sh  -- 2 GPUs: multi-GPU parity tests, sharded emulation tests, stage times, bench line
timeout 600 python -m pytest tests/test_gpu_multi.py tests/test_gpu_sharded.py -q -x 2>&1 | tail -3 | tee gpurun_out/r2_multi_gpu_parity_2.txt
export FOCAL_B200_STAGE_TIMES=1
for B in 8192 2048; do FB_B=$B python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 tools/shard_stage_times.py 2>&1 | grep -E "rank|row-sharded"; done | tee gpurun_out/r2_shard_stage_times_2.txt
unset FOCAL_B200_STAGE_TIMES
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 100 --warmup 5 > gpurun_out/r2_bench_2gpu.json 2> gpurun_out/r2_bench_2gpu.err
python -c "
import json
d=json.loads(open('gpurun_out/r2_bench_2gpu.json').read().strip().splitlines()[-1])
print({k:d.get(k) for k in ('n_gpus','value','ms_per_step','loss')}); print(d['e2e'])"
