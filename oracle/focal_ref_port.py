"""Op-for-op CPU port of the reference FOCAL loss, used as the timed CPU baseline.

TEST / BENCH INFRASTRUCTURE ONLY -- see ``oracle/__init__.py``.

The reference is pure Python on stock PyTorch and cannot travel to the GPU box
(/root/reference does not exist there), so ``bench.py``'s ``cpu_baseline`` and
``--impl reference`` legs time THIS port.  Unlike ``focal_oracle.py`` (closed
form, Gram matrices) it deliberately keeps the reference's cost structure: the same
ATen operators in the same order, including the broadcast cosine similarity that
materialises [S, N, N, d] and the boolean-mask gather, and autograd for the
backward pass.  What it follows:

    loss.py:74      broadcast nn.CosineSimilarity(dim=-1) / T        -> _pairwise_cosine
    loss.py:75-80   +-b diagonals as positives, boolean mask gather  -> _info_nce
    loss.py:83-85   CrossEntropy(mean) with label 0
    loss.py:96-104  CosineEmbeddingLoss(mean), target -1             -> _orthogonality
    loss.py:113-135 cdist, masked block mean, MarginRankingLoss      -> _temporal_rank
    loss.py:152-216 reshape, split, loops, weighted sum              -> focal_loss_port

Pinned against the live reference by ``tests/golden`` (see ``oracle/make_golden.py``).
"""
from __future__ import annotations

from typing import Dict

import torch
import torch.nn.functional as F

from .focal_oracle import FocalConfig


def _negatives_mask(S: int, b: int) -> torch.Tensor:
    """[S, 2b, 2b] bool: False on the diagonal and on the two +-b diagonals (loss.py:35-44)."""
    N = 2 * b
    keep = torch.ones(N, N)
    keep.fill_diagonal_(0)
    eye = torch.eye(b)
    keep[:b, b:] -= eye
    keep[b:, :b] -= eye
    return keep.unsqueeze(0).repeat(S, 1, 1).bool()


def _pairwise_cosine(z: torch.Tensor) -> torch.Tensor:
    # same operator, same broadcast as the reference: materialises [S, N, N, d]
    return F.cosine_similarity(z.unsqueeze(2), z.unsqueeze(1), dim=-1)


def _info_nce(e1: torch.Tensor, e2: torch.Tensor, T: float) -> torch.Tensor:
    """e1, e2: [b, S, w] views."""
    b, S, _ = e1.shape
    N = 2 * b
    z = torch.cat((e1.transpose(0, 1), e2.transpose(0, 1)), dim=1)        # [S, N, w]
    sim = _pairwise_cosine(z) / T
    up = torch.diagonal(sim, b, dim1=-2, dim2=-1)
    lo = torch.diagonal(sim, -b, dim1=-2, dim2=-1)
    pos = torch.cat((up, lo), dim=1).reshape(S, N, 1)
    neg = sim[_negatives_mask(S, b)].reshape(S, N, -1)
    logits = torch.cat((pos, neg), dim=2).reshape(S * N, -1)
    target = torch.zeros(S * N, dtype=torch.long)
    return F.cross_entropy(logits, target, reduction="mean")


def _orthogonality(u: torch.Tensor, v: torch.Tensor) -> torch.Tensor:
    u2 = u.reshape(-1, u.shape[-1])
    v2 = v.reshape(-1, v.shape[-1])
    return F.cosine_embedding_loss(u2, v2, -torch.ones(u2.shape[0], dtype=u2.dtype), reduction="mean")


def _offdiag(m: torch.Tensor) -> torch.Tensor:
    """Row-major off-diagonal entries of a square matrix (tensor_utils.py:199-211)."""
    n = m.shape[0]
    return m.flatten()[1:].view(n - 1, n + 1)[:, :-1].reshape(n, n - 1)


def _temporal_rank(x: torch.Tensor, margin: float) -> torch.Tensor:
    """x: [b, S, D]."""
    b, S, D = x.shape
    flat = x.reshape(b * S, D)
    dist = torch.cdist(flat, flat, p=2).reshape(b, S, b, S).permute(0, 2, 1, 3)
    keep = torch.ones(b * S, b * S, dtype=x.dtype).fill_diagonal_(0)
    keep = keep.reshape(b, S, b, S).permute(0, 2, 1, 3)
    seq = (dist * keep).sum(dim=[2, 3]) / keep.sum(dim=[2, 3])
    intra = torch.diagonal(seq, 0).repeat_interleave(b - 1)
    inter = _offdiag(seq).flatten()
    return F.margin_ranking_loss(intra, inter, -torch.ones_like(intra), margin=margin, reduction="mean")


def focal_loss_port(f1: Dict[str, torch.Tensor], f2: Dict[str, torch.Tensor], cfg: FocalConfig) -> torch.Tensor:
    """Differentiable scalar; call ``.backward()`` on it like pretrain.py:70 does."""
    mods = list(cfg.modalities)
    S = cfg.seq_len
    views = []
    for f in (f1, f2):
        full = {m: f[m].reshape(-1, S, f[m].shape[-1]) for m in mods}
        d = {m: full[m].shape[-1] // 2 for m in mods}
        views.append({
            "full": full,
            "shared": {m: full[m][:, :, : d[m]] for m in mods},
            "private": {m: full[m][:, :, d[m]: 2 * d[m]] for m in mods},
        })

    shared = 0
    key = "full" if cfg.no_private else "shared"
    for vw in views:
        for i, a in enumerate(mods):
            for c in mods[i + 1:]:
                shared = shared + _info_nce(vw[key][a], vw[key][c], cfg.temperature)

    private = 0
    for m in mods:
        private = private + _info_nce(views[0]["private"][m], views[1]["private"][m], cfg.temperature)

    temporal = 0
    for vw in views:
        for m in mods:
            temporal = temporal + _temporal_rank(vw["full"][m], cfg.margin)

    orth = 0
    for vw in views:
        for i, a in enumerate(mods):
            orth = orth + _orthogonality(vw["shared"][a], vw["private"][a])
            for c in mods[i + 1:]:
                orth = orth + _orthogonality(vw["private"][a], vw["private"][c])

    return cfg.w_shared * shared + cfg.w_private * private + cfg.w_orth * orth + cfg.w_rank * temporal
