timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_multi.py -q -x 2>&1 | tail -2
export FOCAL_B200_STAGE_TIMES=1
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 tools/shard_stage_times.py 2>&1 | grep -v "OMP_NUM\|\*\*\*\*" | tee gpurun_out/r2_shard_stage_times_2.txt
