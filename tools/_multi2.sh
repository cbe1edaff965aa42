# usage: bash tools/_multi2.sh  -- 2 GPUs: multi-GPU parity tests, sharded emulation tests, stage times, bench lines (multicast on / off)
timeout 600 python -m pytest tests/test_gpu_multi.py tests/test_gpu_sharded.py -q -x -s 2>&1 | grep -E "multi-GPU parity|passed|failed|skipped|Error" | tee gpurun_out/r2_multi_gpu_parity_2.txt
for mc in 1 0; do
export FOCAL_B200_MULTICAST=$mc
for B in 8192 2048; do echo "multicast=$mc"; FOCAL_B200_STAGE_TIMES=1 FB_B=$B python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 tools/shard_stage_times.py 2>&1 | grep -E "rank|row-sharded"; done
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 100 --warmup 5 > gpurun_out/r2_bench_2gpu_mc$mc.json 2> gpurun_out/r2_bench_2gpu_mc$mc.err
python -c "
import json
d=json.loads(open('gpurun_out/r2_bench_2gpu_mc$mc.json').read().strip().splitlines()[-1])
print('multicast=$mc', {k:d.get(k) for k in ('n_gpus','value','ms_per_step','loss')}); print(d['e2e'])"
done 2>&1 | tee gpurun_out/r2_shard_stage_times_2.txt
tail -5 gpurun_out/r2_bench_2gpu_mc1.err
