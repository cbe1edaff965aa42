"""Frequency-domain input stage of the reference on the B200.

Mirrors, for CUDA tensors, two pieces of the reference that sit right before the backbone:

* ``Augmenter.fft_preprocess`` (/root/reference/src/data_augmenter/Augmenter.py:141-158): FFT along the last axis
  (cuFFT through ``torch.fft.fft`` -- a library transform, like cuBLAS for a plain GEMM), then view_as_real /
  permute / reshape to ``[b, 2c, i, s]``;
* ``PhaseShiftAugmenter.forward`` (/root/reference/src/data_augmenter/PhaseShiftAugmenter.py:20-66): with probability
  ``p`` per (location, modality) rotate every complex bin by ONE random angle.

Everything after the FFT is a single hand-written kernel (``focal_b200_spectrum_rotate``: layout change + rotation in
one pass, 16-byte vectors) instead of the reference's clone / abs / angle / cos / sin / mul / stack / permute passes.
The Python ``random`` stream is consumed exactly like the reference does (one draw per (location, modality), a second
one for the angle when the first is below ``p``), so a seeded run makes the same decisions and angles.

There is no CPU path: CPU tensors raise.
"""
from __future__ import annotations

import ctypes as C
import math
from random import random
from typing import Dict, Optional

import torch
import torch.nn as nn

from . import _cabi


def _rotate(x: torch.Tensor, interleaved: bool, n_bc: int, plane: int, angle: float, out_shape) -> torch.Tensor:
    if not x.is_cuda:
        raise RuntimeError("focal_b200.augment runs on CUDA tensors only (no CPU fallback)")
    if plane % 4:
        raise ValueError(f"intervals x spectrum length must be a multiple of 4, got {plane}")
    lib = _cabi.load()
    out = torch.empty(out_shape, dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        rc = lib.focal_b200_spectrum_rotate(C.c_void_p(x.data_ptr()), C.c_void_p(out.data_ptr()), n_bc, plane,
                                            int(interleaved), math.cos(angle), math.sin(angle),
                                            C.c_void_p(torch.cuda.current_stream(x.device).cuda_stream))
    _cabi.check(rc, "focal_b200_spectrum_rotate")
    return out


def spectrum_to_channels(spec: torch.Tensor, angle: float = 0.0) -> torch.Tensor:
    """complex64 ``[b, c, i, s]`` -> fp32 ``[b, 2c, i, s]`` (channel 2k = real, 2k + 1 = imaginary part of complex
    channel k), every bin rotated by ``angle``: the tail of fft_preprocess (+ a fused phase shift)."""
    if spec.dtype != torch.complex64 or spec.dim() != 4:
        raise TypeError(f"expected a complex64 [b, c, i, s] spectrum, got {spec.dtype} {tuple(spec.shape)}")
    b, c, i, s = spec.shape
    spec = spec.contiguous()
    return _rotate(torch.view_as_real(spec), True, b * c, i * s, angle, (b, 2 * c, i, s))


def fft_preprocess(time_loc_inputs: Dict[str, Dict[str, torch.Tensor]]) -> Dict[str, Dict[str, torch.Tensor]]:
    """Drop-in for ``Augmenter.fft_preprocess``: ``{loc: {mod: [b, c, i, s]}}`` time domain -> frequency domain."""
    out: Dict[str, Dict[str, torch.Tensor]] = {}
    for loc, mods in time_loc_inputs.items():
        out[loc] = {mod: spectrum_to_channels(torch.fft.fft(x.float(), dim=-1)) for mod, x in mods.items()}
    return out


def phase_shift(x: torch.Tensor, angle: float) -> torch.Tensor:
    """``[b, 2c, i, s]`` planar (re, im) channels -> the same, every complex bin rotated by ``angle``."""
    if x.dtype != torch.float32 or x.dim() != 4 or x.shape[1] % 2:
        raise TypeError(f"expected fp32 [b, 2c, i, s], got {x.dtype} {tuple(x.shape)}")
    b, c2, i, s = x.shape
    return _rotate(x.contiguous(), False, b * (c2 // 2), i * s, angle, (b, c2, i, s))


class PhaseShiftAugmenter(nn.Module):
    """Same constructor, ``forward(org_loc_inputs, labels)`` signature, return triple and ``random`` consumption as the
    reference class (PhaseShiftAugmenter.py:9-66)."""

    def __init__(self, args) -> None:
        super().__init__()
        self.args = args
        self.config = args.dataset_config["phase_shift"]
        self.p = self.config["prob"]
        self.modalities = args.dataset_config["modality_names"]
        self.locations = args.dataset_config["location_names"]

    def forward(self, org_loc_inputs, labels):
        aug_loc_inputs, aug_mod_labels = {}, []
        b: Optional[int] = None
        for loc in self.locations:
            aug_loc_inputs[loc] = {}
            for mod in self.modalities:
                if b is None:
                    b = org_loc_inputs[loc][mod].shape[0]
                if random() < self.p:
                    angle = (random() - 0.5) * 2 * math.pi
                    aug_loc_inputs[loc][mod] = phase_shift(org_loc_inputs[loc][mod], angle)
                    aug_mod_labels.append(1)
                else:
                    aug_loc_inputs[loc][mod] = org_loc_inputs[loc][mod]
                    aug_mod_labels.append(0)
        lab = torch.tensor(aug_mod_labels, dtype=torch.float32, device=self.args.device)
        return aug_loc_inputs, lab.unsqueeze(0).tile([b, 1]).float(), labels


def install(augmenter) -> None:
    """Swap the two stages into a live reference ``Augmenter`` instance (data_augmenter is a regular package, so it
    cannot be shadowed per module like ``models.loss``): its ``fft_preprocess`` and every PhaseShiftAugmenter it holds."""
    augmenter.fft_preprocess = fft_preprocess
    for name in ("augmenters", "time_augmenters", "freq_augmenters"):
        lst = getattr(augmenter, name, None)
        if isinstance(lst, list):
            for k, a in enumerate(lst):
                if type(a).__name__ == "PhaseShiftAugmenter" and not isinstance(a, PhaseShiftAugmenter):
                    lst[k] = PhaseShiftAugmenter(a.args)
