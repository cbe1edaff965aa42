"""GPU, >= 2 devices: the row-sharded loss on REAL peers (one process per GPU, NCCL for the handshake only) against the
fp64 oracle on the whole batch -- loss on every rank, gradients of every rank's own rows (VERDICT r1 missing #8).

All exchange paths are covered: kernel stores into NVLink peer memory (`focal_b200_loss_sharded`, the default), the same
with NVSwitch multicast stores (workspaces from torch symmetric memory), and the collective path (NCCL all-gathers).  Skips when fewer than two GPUs are visible (the single-GPU box emulates the ranks on
streams instead: tests/test_gpu_sharded.py).  `tools/dist_gpu_check.py` prints the same comparison for profiles/.
"""
import os
import socket

import pytest
import torch

pytestmark = pytest.mark.gpu

CASES = [
    # B, D, mods, T, precision
    (2048, 256, ("seismic", "audio"), 0.5, "bf16"),
    (2048, 256, ("seismic", "audio"), 0.5, "fp32"),
    (1536, 128, ("acc", "gyr", "mag"), 0.07, "fp32"),
    (8192, 256, ("seismic", "audio"), 0.5, "bf16"),          # BASELINE.json metric configuration
]


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    import torch.distributed as dist

    from focal_b200.engine import FocalEngine, FocalHyper
    from oracle import focal_oracle as fo
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    res = []
    try:
        for (B, D, mods, T, prec) in CASES:
            S = 4
            if B % (S * world):
                continue
            f1, f2 = fo.make_structured(5, list(mods), B, D, S)
            cfg = fo.FocalConfig(modalities=list(mods), seq_len=S, temperature=T)
            ref = fo.focal_closed_form({m: v.to(dev) for m, v in f1.items()}, {m: v.to(dev) for m, v in f2.items()}, cfg,
                                       dtype=torch.float64)
            want = [ref.grads1[m] for m in mods] + [ref.grads2[m] for m in mods]
            hp = FocalHyper(tuple(mods), S, T, 1.0, 1.0, 1.0, 3.0, 5.0, False, 7, prec)
            Bl = B // world
            l1 = {m: v[rank * Bl:(rank + 1) * Bl].to(dev) for m, v in f1.items()}
            l2 = {m: v[rank * Bl:(rank + 1) * Bl].to(dev) for m, v in f2.items()}
            for mode in ("peer", "peer_mc", "collective"):
                os.environ["FOCAL_B200_PEER"] = "0" if mode == "collective" else "1"
                # per-peer stores / NVSwitch multicast stores (multimem.st; falls back to per-peer stores on a fabric
                # without multicast support)
                os.environ["FOCAL_B200_MULTICAST"] = "1" if mode == "peer_mc" else "0"
                eng = FocalEngine(hp, process_group=dist.group.WORLD)
                for _ in range(4):                  # eager, capture, replays: the barrier epochs must stay in step
                    loss5, grads = eng.loss_and_grads(l1, l2, True)
                torch.cuda.synchronize()
                lerr = float((loss5[0].double() - ref.loss).abs() / ref.loss.abs())
                gerr = max(float((g.double() - w[rank * Bl:(rank + 1) * Bl]).norm() / w[rank * Bl:(rank + 1) * Bl].norm())
                           for g, w in zip(grads, want))
                res.append((B, D, len(mods), prec, mode, lerr, gerr, eng.graph_replays))
                dist.barrier()
                eng.close()
                dist.barrier()
        out[rank] = res
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(900)
def test_row_sharded_loss_on_real_peers_matches_the_oracle():
    assert torch.cuda.is_available(), "GPU tests selected (-m gpu) but no CUDA device is visible"
    world = min(torch.cuda.device_count(), 8)
    if world < 2:
        pytest.skip("needs at least two GPUs (run under `gpurun --gpus 2` or 8)")
    while 2048 % (4 * world):
        world -= 1
    import torch.multiprocessing as mp
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    assert len(out) == world
    for rank in range(world):
        assert out[rank], f"rank {rank} ran no case"
        for (B, D, M, prec, mode, lerr, gerr, replays) in out[rank]:
            what = (rank, world, B, D, M, prec, mode)
            assert lerr < 1e-4, (what, "loss", lerr)
            assert gerr < (2e-3 if prec == "fp32" else 1e-2), (what, "grad", gerr)
            if mode != "collective":
                assert replays >= 2, (what, "the peer path must replay its captured step")
    # a one-line record for profiles/
    worst_l = max(r[5] for rank in range(world) for r in out[rank])
    worst_g = {p: max(r[6] for rank in range(world) for r in out[rank] if r[3] == p) for p in ("bf16", "fp32")}
    print(f"\nmulti-GPU parity, {world} ranks: worst loss rel err {worst_l:.2e}; worst grad rel err "
          f"bf16 {worst_g['bf16']:.2e}, fp32 {worst_g['fp32']:.2e}")
