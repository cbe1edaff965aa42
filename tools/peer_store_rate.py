"""SM-store rate into NVLink peer memory (torchrun, >= 2 GPUs): every rank pushes a slice of `MB` MiB into the symmetric
buffers of all ranks at once (the traffic pattern of the row-sharded prologue) -- plain 16-byte stores per destination
against one NVSwitch multicast store (multimem.st) -- for several grid sizes.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 \
        tools/peer_store_rate.py
"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
import torch.distributed._symmetric_memory as symm

from focal_b200._cabi import load_bringup


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", device_id=dev)
    lib = load_bringup()
    lib.focal_b200_debug_peer_store.argtypes = [C.c_void_p, C.c_int, C.c_uint32, C.c_uint32, C.c_int, C.c_int, C.c_void_p]
    lib.focal_b200_debug_peer_store.restype = C.c_int
    slice_mb = float(os.environ.get("MB", 2))
    nbytes = int(slice_mb * 2 ** 20)
    buf = symm.empty(world * nbytes, dtype=torch.uint8, device=dev)
    h = symm.rendezvous(buf, dist.group.WORLD.group_name)
    buf.zero_()
    peers = (C.c_void_p * 8)(*[int(p) for p in h.buffer_ptrs])
    mc = (C.c_void_p * 8)(int(h.multicast_ptr))
    st = torch.cuda.current_stream().cuda_stream
    rows = []
    for mode, name, dst, n in ((0, "unicast chunk-major", peers, world), (2, "unicast dest-major", peers, world),
                               (1, "multimem.st", mc, 1)):
        if mode == 1 and not h.multicast_ptr:
            continue
        for grid in (32, 64, 148, 296, 592):
            for everyone in (True, False):
                dist.barrier()
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                reps = 20
                active = everyone or rank == 0
                for k in range(reps + 3):
                    if k == 3:
                        e0.record()
                    if active:
                        rc = lib.focal_b200_debug_peer_store(dst, n, rank * nbytes, nbytes, mode, grid, st)
                        assert rc == 0, rc
                e1.record()
                torch.cuda.synchronize()
                t = e0.elapsed_time(e1) / reps * 1e3
                rows.append((name, grid, "all ranks" if everyone else "rank 0 only", t))
    # check the multicast landed everywhere: word 0 of each rank's slice holds threadIdx 0 / block 0 pattern (zeros),
    # word 2 holds 1.0f
    torch.cuda.synchronize()
    dist.barrier()
    ok = all(int(buf.view(torch.int32)[r * nbytes // 4 + 2].item()) == 0x3f800000 for r in range(world))
    allr = [None] * world
    dist.all_gather_object(allr, (rows, ok))
    if rank == 0:
        print(f"{world} GPUs, slice of {slice_mb} MiB per rank -> every rank's buffer; us per launch (rank 0 / max over ranks); all slices landed: "
              f"{all(o for _, o in allr)}")
        for i, (name, grid, who, t) in enumerate(rows):
            tm = max(a[0][i][3] for a in allr) if who == "all ranks" else t
            egress = nbytes * (world - 1 if name != "multimem.st" else 1)
            print(f"  {name:20s} grid {grid:4d} {who:12s}: {t:7.1f} / {tm:7.1f} us   ({nbytes * (world - 1) / tm / 1e3:7.1f} GB/s delivered to peers per rank, "
                  f"{egress / tm / 1e3:7.1f} GB/s egress)")
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
