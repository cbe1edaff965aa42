"""Multi-GPU parity check (run under torchrun on N GPUs): the row-sharded loss == the single-GPU loss.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        tools/dist_gpu_check.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

from focal_b200.engine import FocalEngine, FocalHyper
from oracle import focal_oracle as fo


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
    dist.init_process_group("nccl", device_id=dev)
    ok = True
    for (B, D, mods, T) in ((2048, 256, ("seismic", "audio"), 0.5), (1536, 128, ("acc", "gyr", "mag"), 0.07),
                            (8192, 256, ("seismic", "audio"), 0.5)):
        S = 4
        f1, f2 = fo.make_structured(5, mods, B, D, S)
        hp = FocalHyper(tuple(mods), S, T, 1.0, 1.0, 1.0, 3.0, 5.0)
        single = FocalEngine(hp)                                    # whole batch on this GPU
        full5, fullg = single.loss_and_grads({m: v.to(dev) for m, v in f1.items()},
                                             {m: v.to(dev) for m, v in f2.items()}, True)
        full5, fullg = full5.clone(), [g.clone() for g in fullg]
        Bl = B // world
        l1 = {m: v[rank * Bl:(rank + 1) * Bl].to(dev) for m, v in f1.items()}
        l2 = {m: v[rank * Bl:(rank + 1) * Bl].to(dev) for m, v in f2.items()}
        for mode in ("peer", "collective"):
            os.environ["FOCAL_B200_PEER"] = "1" if mode == "peer" else "0"
            sharded = FocalEngine(hp, process_group=dist.group.WORLD)
            for _ in range(4):                      # eager, capture, replays: barrier epochs must stay in step
                loss5, grads = sharded.loss_and_grads(l1, l2, True)
            torch.cuda.synchronize()
            # total loss relative; the sub-loss sums are compared with an absolute floor (at T = 0.07 the InfoNCE parts
            # are ~1e-3 differences of ~13-sized fp32 sums, so their last bits depend on the per-rank summation split)
            lerr = max(float((loss5[0] - full5[0]).abs() / full5[0].abs()),
                       float(((loss5[1:] - full5[1:]).abs() / full5[1:].abs().clamp_min(1.0)).max()))
            gerr = max(float((g - fg[rank * Bl:(rank + 1) * Bl]).norm() / fg[rank * Bl:(rank + 1) * Bl].norm())
                       for g, fg in zip(grads, fullg))
            good = lerr < 2e-6 and gerr < 1e-5
            ok = ok and good
            print(f"[rank {rank}/{world}] {mode:10s} B={B} D={D} M={len(mods)}: loss5 rel diff {lerr:.2e}, "
                  f"grad rel diff {gerr:.2e} {'OK' if good else 'MISMATCH'}", flush=True)
            dist.barrier()
            sharded.close()
            dist.barrier()
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
