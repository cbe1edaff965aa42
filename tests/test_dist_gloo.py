"""CPU, world size 2, gloo: the row-sharded multi-GPU host logic.

The engine's collectives (all-gather of features, all-gather of InfoNCE row sums, all-reduce of the loss partials) run
for real over gloo; the per-rank numerical work is done by a TEST-ONLY backend built on the oracle (the product backend
needs a GPU and is exercised by tests/test_gpu_parity.py::test_row_shards_on_one_gpu_sum_to_global).
"""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from focal_b200.engine import FocalEngine, FocalHyper
from oracle import focal_oracle as fo

MODS = ("seismic", "audio")
B, D, S = 96, 32, 4


class OracleBackend:
    """Same contract as focal_b200.engine.CudaBackend.run, evaluated with the CPU oracle (tests only)."""
    name = "oracle-test-double"

    def run(self, hp, feats, seq, need_grad, exchange_rowsum=None):
        M = len(hp.modalities)
        cfg = fo.FocalConfig(modalities=list(hp.modalities), seq_len=hp.seq_len, temperature=hp.temperature,
                             margin=hp.margin, w_shared=hp.w_shared, w_private=hp.w_private, w_orth=hp.w_orth,
                             w_rank=hp.w_rank, no_private=hp.no_private)
        f1 = {m: feats[i] for i, m in enumerate(hp.modalities)}
        f2 = {m: feats[M + i] for i, m in enumerate(hp.modalities)}
        res = fo.focal_closed_form(f1, f2, cfg, dtype=torch.float64, need_grad=need_grad, seq_rows=seq)
        if exchange_rowsum is not None and need_grad:
            # the exchange must deliver every other rank's row sums: fill only the owned slice, poison the rest
            b = feats[0].shape[0] // hp.seq_len
            full = torch.stack([torch.stack((r[:, :b], r[:, b:]), dim=1) for r in res.aux["nce_rowsum"]]).float()
            rs = torch.full_like(full, float("nan"))
            rs[..., seq[0]:seq[1]] = full[..., seq[0]:seq[1]]
            exchange_rowsum(rs)
            assert torch.allclose(rs, full, rtol=1e-6), "row-sum exchange did not deliver the other ranks' rows"
        loss5 = torch.stack([res.loss] + [res.parts[k] for k in ("shared", "private", "orth", "temporal")]).float()
        grads = None
        if need_grad:
            grads = [res.grads1[m].float() for m in hp.modalities] + [res.grads2[m].float() for m in hp.modalities]
        return loss5, grads


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.set_num_threads(1)
        f1, f2 = fo.make_structured(21, MODS, B, D, S)
        hp = FocalHyper(MODS, S, 0.5, 1.0, 1.0, 1.0, 3.0, 5.0)
        eng = FocalEngine(hp, process_group=dist.group.WORLD, backend=OracleBackend())
        Bl = B // world
        l1 = {m: v[rank * Bl:(rank + 1) * Bl] for m, v in f1.items()}
        l2 = {m: v[rank * Bl:(rank + 1) * Bl] for m, v in f2.items()}
        loss5, grads = eng.loss_and_grads(l1, l2, True)
        loss5_ng, grads_ng = eng.loss_and_grads(l1, l2, False)
        out[rank] = (loss5.clone(), [g.clone() for g in grads], loss5_ng.clone(), grads_ng)
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_rank_row_shards_match_single_process():
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    f1, f2 = fo.make_structured(21, MODS, B, D, S)
    cfg = fo.FocalConfig(modalities=list(MODS), seq_len=S, temperature=0.5)
    ref = fo.focal_closed_form(f1, f2, cfg, dtype=torch.float64)
    want5 = torch.stack([ref.loss] + [ref.parts[k] for k in ("shared", "private", "orth", "temporal")]).float()
    Bl = B // world
    for rank in range(world):
        loss5, grads, loss5_ng, grads_ng = out[rank]
        assert torch.allclose(loss5, want5, rtol=1e-5), (rank, loss5, want5)          # every rank holds the global loss
        assert torch.allclose(loss5_ng, want5, rtol=1e-5) and grads_ng is None
        want = [ref.grads1[m] for m in MODS] + [ref.grads2[m] for m in MODS]
        for g, w in zip(grads, want):
            assert g.shape == (Bl, D)
            assert torch.allclose(g.double(), w[rank * Bl:(rank + 1) * Bl], rtol=1e-4, atol=1e-7)


def test_single_process_engine_with_test_backend():
    f1, f2 = fo.make_iid(3, MODS, 64, 16)
    hp = FocalHyper(MODS, S, 0.5, 1.0, 1.0, 1.0, 3.0, 5.0)
    eng = FocalEngine(hp, backend=OracleBackend())
    loss5, grads = eng.loss_and_grads(f1, f2, True)
    ref = fo.focal_closed_form(f1, f2, fo.FocalConfig(modalities=list(MODS), seq_len=S), dtype=torch.float64)
    assert float(loss5[0]) == pytest.approx(float(ref.loss), rel=1e-6)
    assert len(grads) == 4 and grads[0].shape == (64, 16)
