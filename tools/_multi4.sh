# 4 GPUs: parity test on real peers + headline bench line
timeout 600 python -m pytest tests/test_gpu_multi.py -q -s 2>&1 | grep -E "multi-GPU parity|passed|failed|skipped|Error" | tee gpurun_out/r2_multi_gpu_parity_4.txt
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 4 --steps 100 --warmup 5 > gpurun_out/r2_bench_4gpu.json 2> gpurun_out/r2_bench_4gpu.err
python -c "
import json
d=json.loads(open('gpurun_out/r2_bench_4gpu.json').read().strip().splitlines()[-1])
print({k:d.get(k) for k in ('n_gpus','value','ms_per_step','loss')}); print(d['e2e'])"
