"""Per-stage device times of the row-sharded step on real GPUs (run under torchrun, FOCAL_B200_STAGE_TIMES=1): every rank
enqueues the whole launch sequence behind a spin kernel (so host launch latency is not in the figures) right after a
barrier, CUDA events between the launches give the time of each stage INCLUDING what it waits for from the peers.

    FOCAL_B200_STAGE_TIMES=1 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 \
        --master-port 29517 tools/shard_stage_times.py
"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

from focal_b200.engine import FocalEngine, FocalHyper


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", device_id=dev)
    B, D, S, mods = int(os.environ.get("FB_B", 8192)), 256, 4, ("seismic", "audio")
    hp = FocalHyper(mods, S, 0.5, 1.0, 1.0, 1.0, 3.0, 5.0)
    eng = FocalEngine(hp, process_group=dist.group.WORLD, use_cuda_graph=False)
    Bl = B // world
    torch.manual_seed(rank)
    l1 = {m: torch.randn(Bl, D, device=dev) for m in mods}
    l2 = {m: torch.randn(Bl, D, device=dev) for m in mods}
    lib = eng.backend.lib
    lib.focal_b200_debug_stage_times.argtypes = [C.c_void_p, C.c_int]
    lib.focal_b200_debug_stage_times.restype = C.c_int
    names = ["prologue", "nce_rowsum", "nce_lse", "temporal", "nce_grad", "finalize"]
    acc = [0.0] * 6
    steps, warm = 30, 5
    for k in range(steps + warm):
        dist.barrier()
        torch.cuda.synchronize()
        torch.cuda._sleep(3_000_000)
        eng.loss_and_grads(l1, l2, True)
        buf = (C.c_float * 8)()
        n = lib.focal_b200_debug_stage_times(buf, 8)
        assert n == 6, n
        if k >= warm:
            for i in range(6):
                acc[i] += buf[i] / steps
    tot = sum(acc)
    allv = [None] * world
    dist.all_gather_object(allv, acc)
    if rank == 0:
        print(f"row-sharded step, {world} GPUs, B={B}: mean device time per stage in us (incl. waits for the peers)")
        for r, a in enumerate(allv):
            print(f"  rank {r}: " + " ".join(f"{n}={v * 1e3:6.1f}" for n, v in zip(names, a)) + f" | sum {sum(a) * 1e3:6.1f}")
    dist.barrier()
    eng.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
