"""Bring-up diagnostics (GPU): per-term loss / gradient errors of the CUDA path against the fp64 oracle."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import focal_oracle as fo
from tests._golden import config_of, load_case
from focal_b200.engine import CudaBackend, FocalHyper
import dataclasses

names = sys.argv[1:] or ["kat4_m4"]
be = CudaBackend()
for name in names:
    case, rec, f1, f2 = load_case(name)
    cfg = config_of(case)
    mods = list(cfg.modalities)
    S = cfg.seq_len
    feats = [f1[m].cuda() for m in mods] + [f2[m].cuda() for m in mods]
    B, D = feats[0].shape
    b = B // S
    for tname, mask, w in (("nce", 1, dict(w_orth=0.0, w_rank=0.0)), ("orth", 2, dict(w_shared=0.0, w_private=0.0, w_rank=0.0)),
                           ("temporal", 4, dict(w_shared=0.0, w_private=0.0, w_orth=0.0)), ("all", 7, {})):
        c2 = dataclasses.replace(cfg, **w)
        ref = fo.focal_closed_form(f1, f2, c2, dtype=torch.float64)
        hp = FocalHyper(tuple(mods), S, cfg.temperature, cfg.margin, cfg.w_shared, cfg.w_private, cfg.w_orth,
                        cfg.w_rank, cfg.no_private, mask)
        if b <= 1 or S <= 1:
            if mask & 4 and mask != 7:
                continue
        loss5, grads = be.run(hp, feats, (0, b), True, None)
        torch.cuda.synchronize()
        l5 = loss5.cpu().double()
        refparts = [float(ref.parts[k]) for k in ("shared", "private", "orth", "temporal")]
        print(f"[{name}] term={tname:8s} loss5={[round(float(v), 6) for v in l5]} ref_parts={[round(v, 6) for v in refparts]}")
        rg = [ref.grads1[m] for m in mods] + [ref.grads2[m] for m in mods]
        if tname == "temporal" and b > 1 and S > 1:
            ws, info = be.workspace(be._cfg(hp, B, D, True, (0, b)), feats[0].device)
            cnt = be._view(ws, info.cnt_off, info.cnt_bytes, torch.int32, (len(feats), info.bpad)).cpu()
            mi = be._view(ws, info.mintra_off, len(feats) * info.Bpad * 4, torch.float32, (len(feats), info.Bpad)).cpu()
            for t in range(len(feats)):
                ax = ref.aux["temporal"][t]
                want = ax["active"].sum(dim=1)
                got = cnt[t, :b].long()
                h = ax["mII"][:, None] - ax["m"] + cfg.margin
                h = h + torch.eye(b, dtype=h.dtype) * 1e9
                bad = (got != want).nonzero().flatten().tolist()
                mierr = float((mi[t, :B:S].double() - ax["mII"]).abs().max())
                print(f"    t={t} cnt mismatches at I={bad[:8]} got={got[bad[:8]].tolist()} want={want[bad[:8]].tolist()} "
                      f"min|h| per bad row={[float(h[i].abs().min()) for i in bad[:8]]} global min|h|={float(h.abs().min()):.2e} mII err={mierr:.2e}")
        for t, (g, r) in enumerate(zip(grads, rg)):
            g = g.cpu().double()
            err = (g - r).norm() / r.norm().clamp_min(1e-300)
            rowerr = (g - r).norm(dim=1)
            worst = int(rowerr.argmax())
            flag = "  <<<<" if err > 1e-2 else ""
            print(f"    t={t} rel={float(err):.3e} worst_row={worst} rowerr={float(rowerr[worst]):.3e} "
                  f"rownorm={float(r[worst].norm()):.3e} nbad_rows={int((rowerr > 0.05 * r.norm(dim=1).clamp_min(1e-30)).sum())}{flag}")
