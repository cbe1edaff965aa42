# usage: bash tools/_dbg2.sh N  -- prologue store experiments on N GPUs (timing only)
N=$1
export FOCAL_B200_STAGE_TIMES=1
for B in 8192 $((1024*N)); do for v in "0 0" "0 1" "3 0" "7 0" "16 0" "23 0"; do
set -- $v
echo "== B=$B dbg=$1 replicas=$2"
FB_B=$B FOCAL_B200_PROLOGUE_DBG=$1 FOCAL_B200_PROLOGUE_REPLICAS=$2 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 tools/shard_stage_times.py 2>&1 | grep -E "rank 0"
done; done | tee gpurun_out/r2_prologue_dbg_$N.txt
