export FOCAL_B200_STAGE_TIMES=1
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517 tools/shard_stage_times.py 2>&1 | grep -v "OMP_NUM\|\*\*\*\*" | tee gpurun_out/r2_shard_stage_times_8.txt
