# 8 GPUs: headline bench line with NVSwitch multicast stores on / off
for mc in 1 0; do
FOCAL_B200_MULTICAST=$mc python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 8 --steps 100 --warmup 5 > gpurun_out/r2_bench_8gpu_mc$mc.json 2> gpurun_out/r2_bench_8gpu_mc$mc.err
python -c "
import json
d=json.loads(open('gpurun_out/r2_bench_8gpu_mc$mc.json').read().strip().splitlines()[-1])
print('multicast=$mc', {k:d.get(k) for k in ('n_gpus','value','ms_per_step','loss')}); print(d['e2e'])"
done 2>&1 | tee gpurun_out/r2_mc8.txt
