"""``FOCALLoss`` -- drop-in for the reference module /root/reference/src/models/loss.py:8-218.

Same constructor (``FOCALLoss(args)``), same ``forward(mod_features1, mod_features2, index=None)``, same config
keys (``args.dataset_config["FOCAL"]``: temperature (dict by model or scalar), inter_rank_margin and the four
loss weights; ``["modality_names"]``, ``["seq_len"]``; ``args.model``, ``args.tag``), no parameters and no buffers
(``state_dict()`` is empty, like the reference's).  The arithmetic runs in hand-written sm_100a kernels through the
C ABI of ``libfocal_b200.so``; the gradients w.r.t. every feature tensor are produced in the same launch sequence
as the loss and handed to autograd by a ``torch.autograd.Function``.

Build-side options (never required): ``args.focal_process_group`` -- a ``torch.distributed`` process group over
which the batch is row-sharded (each rank passes its own rows); ``args.focal_precision`` -- ``"fp32"`` (north_star's
fp32 mode: split-bf16 tiles with 16 significant bits, gradients within 2e-3 of the fp32 reference; ``"tf32"`` is accepted
as an alias), ``"bf16"`` (bf16 tiles, 1e-2) or ``"auto"`` (default: fp32 mode for batches up to
``engine.AUTO_FP32_MAX_ROWS`` rows, bf16 above; env ``FOCAL_B200_PRECISION``).

Limits the reference does not have (they raise ``ValueError`` / ``TypeError`` before any launch): see INTEGRATION.md.
Features that are not fp32 (autocast) are upcast, and their gradients cast back by autograd.
"""
from __future__ import annotations

from typing import Dict, Optional

import torch
import torch.nn as nn

from . import _cabi
from .engine import FocalEngine, FocalHyper


class _FocalLossFn(torch.autograd.Function):
    """loss = f(features); gradients are computed together with the loss and saved for backward."""

    @staticmethod
    def forward(ctx, module: "FOCALLoss", n_mod: int, *feats: torch.Tensor):
        mods = module.modalities
        f1 = {m: feats[i] for i, m in enumerate(mods)}
        f2 = {m: feats[n_mod + i] for i, m in enumerate(mods)}
        need_grad = any(ctx.needs_input_grad[2:])
        loss5, grads = module.engine.loss_and_grads(f1, f2, need_grad)
        module.last_parts = loss5.detach()
        if need_grad:
            ctx.save_for_backward(*grads)
        return loss5[0].clone()

    @staticmethod
    def backward(ctx, grad_out: torch.Tensor):
        grads = ctx.saved_tensors
        needs = ctx.needs_input_grad[2:]
        # grad_out is a 0-dim device tensor: scale on the device, no host sync; one multi-tensor launch for all inputs
        wanted = [g for g, need in zip(grads, needs) if need]
        scaled = iter(torch._foreach_mul(wanted, grad_out)) if wanted else iter(())
        return (None, None) + tuple(next(scaled) if need else None for need in needs)


class FOCALLoss(nn.Module):
    def __init__(self, args):
        super().__init__()
        self.args = args
        self.config = args.dataset_config["FOCAL"]
        self.modalities = list(args.dataset_config["modality_names"])
        t = self.config["temperature"]
        self.temperature = float(t[args.model]) if isinstance(t, dict) else float(t)
        if len(self.modalities) > _cabi.FOCAL_MAX_MODALITIES:
            raise ValueError(f"focal_b200 supports up to {_cabi.FOCAL_MAX_MODALITIES} modalities, "
                             f"got {len(self.modalities)}")
        self._engine: Optional[FocalEngine] = None
        self.last_parts: Optional[torch.Tensor] = None   # [total, shared, private, orth, temporal] of the last call

    # the engine (and with it the extension) is created on first use so that constructing the module on a
    # machine without the built library fails at the first forward, loudly, rather than silently falling back
    @property
    def engine(self) -> FocalEngine:
        if self._engine is None:
            cfg = self.config
            hp = FocalHyper(
                modalities=tuple(self.modalities), seq_len=int(self.args.dataset_config["seq_len"]),
                temperature=self.temperature, margin=float(cfg["inter_rank_margin"]),
                w_shared=float(cfg["shared_contrastive_loss_weight"]),
                w_private=float(cfg["private_contrastive_loss_weight"]),
                w_orth=float(cfg["orthogonal_loss_weight"]), w_rank=float(cfg["rank_loss_weight"]),
                no_private=(getattr(self.args, "tag", None) == "noPrivate"),
                precision=str(getattr(self.args, "focal_precision", "auto")).lower())
            self._engine = FocalEngine(hp, process_group=getattr(self.args, "focal_process_group", None))
        return self._engine

    def forward(self, mod_features1: Dict[str, torch.Tensor], mod_features2: Dict[str, torch.Tensor], index=None):
        """loss = w_s * shared InfoNCE + w_p * private InfoNCE + w_o * orthogonality + w_r * temporal ranking."""
        feats = [mod_features1[m] for m in self.modalities] + [mod_features2[m] for m in self.modalities]
        # the reference computes in fp32 (loss.py:74,117); lower-precision activations (autocast) are upcast here so
        # that autograd casts their gradients back
        feats = [f if f.dtype == torch.float32 or not f.is_floating_point() else f.float() for f in feats]
        return _FocalLossFn.apply(self, len(self.modalities), *feats)
