"""Parity margins on the GPU: loss and gradient errors of the CUDA path against the fp64 closed-form oracle
(oracle/focal_oracle.py, itself pinned to the live reference by tests/golden) at the bench sizes.

    python tools/accuracy_report.py > profiles/r2_accuracy.txt
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from focal_b200.engine import CudaBackend, FocalHyper
from oracle import focal_oracle as fo


def main():
    from focal_b200 import _cabi
    print("mode  workload                                   input       loss rel err   grad rel err per tensor (||g - g_ref|| / ||g_ref||)")
    shapes = ((8192, 256, ("seismic", "audio"), 0.5, 4), (4096, 256, ("acc", "gyr", "mag"), 0.07, 4),
              (1024, 256, ("seismic", "audio"), 0.5, 4), (2048, 128, ("seismic", "audio"), 0.5, 4),
              (128, 128, ("seismic", "audio"), 0.5, 4))
    modes = [m for m in os.environ.get("FB_MODES", "fp32,bf16").split(",") if m]
    if os.environ.get("FB_SHAPES"):
        shapes = shapes[:int(os.environ["FB_SHAPES"])]
    for prec, pname in ((_cabi.FOCAL_PREC_FP32, "fp32"), (_cabi.FOCAL_PREC_BF16, "bf16")):
        if pname not in modes:
            continue
        be = CudaBackend(precision=prec)
        report(be, pname, shapes)
    print("tolerances (BASELINE.json north_star): loss 1e-4; gradients 2e-3 (fp32 mode: split-bf16 tiles) / 1e-2 (bf16 mode)")


def report(be, pname, shapes):
    for (B, D, mods, T, S) in shapes:
        for gen in ("iid", "structured"):
            f1, f2 = fo.make_iid(0, list(mods), B, D) if gen == "iid" else fo.make_structured(0, list(mods), B, D, S)
            cfg = fo.FocalConfig(modalities=list(mods), seq_len=S, temperature=T)
            hp = FocalHyper(tuple(mods), S, T, 1.0, 1.0, 1.0, 3.0, 5.0)
            feats = [f1[m].cuda() for m in mods] + [f2[m].cuda() for m in mods]
            loss5, grads = be.run(hp, feats, (0, B // S), True, None)
            torch.cuda.synchronize()
            ref = fo.focal_closed_form({m: v.cuda() for m, v in f1.items()}, {m: v.cuda() for m, v in f2.items()}, cfg,
                                       dtype=torch.float64)
            rg = [ref.grads1[m] for m in mods] + [ref.grads2[m] for m in mods]
            lerr = abs(float(loss5[0]) - float(ref.loss)) / abs(float(ref.loss))
            gerr = [float((g.double() - r.cuda()).norm() / r.cuda().norm()) for g, r in zip(grads, rg)]
            perr = [abs(float(loss5[1 + k]) - float(ref.parts[n])) / max(abs(float(ref.parts[n])), 1e-30)
                    for k, n in enumerate(("shared", "private", "orth", "temporal"))]
            print(f"{pname}  B={B:5d} M={len(mods)} S={S} D={D:3d} T={T:<4}  {gen:10s}  {lerr:.2e}       "
                  + " ".join(f"{e:.2e}" for e in gerr) + "   | parts (shared private orth temporal): "
                  + " ".join(f"{e:.1e}" for e in perr), flush=True)


if __name__ == "__main__":
    main()
