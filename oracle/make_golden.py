"""Generate tests/golden/*.npz from the LIVE, UNMODIFIED reference (run in the build container only).

    python oracle/make_golden.py            # needs /root/reference/src (read-only import)

The reference has no tests or golden vectors of its own, so parity is pinned
to outputs of the reference itself: this script imports ``models.loss.FOCALLoss`` from
/root/reference/src, runs it on seeded synthetic inputs on CPU in fp32 (and fp64 for the
same fp32 draw cast up), and stores loss, the four un-weighted sub-loss sums, and all
gradients.  /root/reference does not exist on the GPU box, so tests only read the fixtures.
"""
from __future__ import annotations

import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
REF_SRC = "/root/reference/src"

from oracle.focal_oracle import make_iid, make_structured  # noqa: E402

DEFAULT_T = {"SW_Transformer": 0.07, "DeepSense": 0.5}

# name, generator, seed, mods, B, D, model, tag, seq_len, temperature override, mutate
CASES = [
    dict(name="kat1_cfg1", gen="iid", seed=0, mods=["seismic", "audio"], B=128, D=128, model="DeepSense"),
    dict(name="kat2", gen="iid", seed=1, mods=["seismic", "audio"], B=256, D=256, model="DeepSense"),
    dict(name="kat3_m3_t007", gen="iid", seed=2, mods=["acc", "gyr", "mag"], B=128, D=256, model="SW_Transformer"),
    dict(name="kat4_m4", gen="iid", seed=3, mods=["m0", "m1", "m2", "m3"], B=64, D=64, model="DeepSense"),
    dict(name="kat5_noprivate", gen="iid", seed=4, mods=["seismic", "audio"], B=64, D=64, model="DeepSense",
         tag="noPrivate"),
    dict(name="skat1", gen="structured", seed=0, mods=["seismic", "audio"], B=128, D=128, model="DeepSense"),
    dict(name="skat2_m3_t007", gen="structured", seed=1, mods=["acc", "gyr", "mag"], B=128, D=256,
         model="SW_Transformer"),
    dict(name="skat3", gen="structured", seed=2, mods=["seismic", "audio"], B=512, D=256, model="DeepSense"),
    # edge cases of SURVEY.md Appendix E
    dict(name="edge_odd_d", gen="iid", seed=5, mods=["seismic", "audio"], B=32, D=33, model="DeepSense"),
    dict(name="edge_zero_row", gen="iid", seed=6, mods=["seismic", "audio"], B=32, D=32, model="DeepSense",
         mutate="zero_row"),
    dict(name="edge_dup_rows", gen="structured", seed=7, mods=["seismic", "audio"], B=48, D=64, model="DeepSense",
         mutate="dup_rows"),
    dict(name="edge_scalar_temp", gen="iid", seed=8, mods=["seismic", "audio"], B=40, D=48, model="DeepSense",
         temperature=0.2),
    dict(name="edge_b1_nan", gen="iid", seed=9, mods=["seismic", "audio"], B=4, D=16, model="DeepSense"),
    dict(name="edge_seq2", gen="iid", seed=10, mods=["seismic", "audio"], B=24, D=32, model="DeepSense", seq_len=2),
    dict(name="edge_seq1_nan", gen="iid", seed=11, mods=["seismic", "audio"], B=16, D=32, model="DeepSense",
         seq_len=1),
    dict(name="edge_ragged_b", gen="structured", seed=12, mods=["seismic", "audio"], B=4 * 37, D=96,
         model="SW_Transformer"),
    # cfg 5 embedding width (shared 256 + private 256): the wide temporal mode of the CUDA path
    dict(name="kat6_d512", gen="structured", seed=13, mods=["seismic", "audio"], B=128, D=512, model="DeepSense"),
    dict(name="kat7_d320_m3", gen="iid", seed=14, mods=["acc", "gyr", "mag"], B=4 * 21, D=320, model="DeepSense"),
    # shapes the reference accepts and round 1 refused: more than 4 modalities, sequence lengths that are not powers of two
    dict(name="kat8_m5", gen="iid", seed=15, mods=["m0", "m1", "m2", "m3", "m4"], B=64, D=64, model="DeepSense"),
    dict(name="kat9_m8", gen="structured", seed=16, mods=[f"m{i}" for i in range(8)], B=32, D=64, model="DeepSense"),
    dict(name="edge_seq3", gen="structured", seed=17, mods=["seismic", "audio"], B=3 * 40, D=128, model="DeepSense",
         seq_len=3),
    dict(name="edge_seq6", gen="iid", seed=18, mods=["seismic", "audio"], B=6 * 24, D=64, model="DeepSense", seq_len=6),
    dict(name="edge_seq5_m3", gen="structured", seed=19, mods=["acc", "gyr", "mag"], B=5 * 32, D=96,
         model="SW_Transformer", seq_len=5),
]


def mutate_inputs(kind, f1, f2, S):
    if kind == "zero_row":
        m0 = next(iter(f1))
        f1[m0][5].zero_()                      # whole row zero: NCE norm clamp + orth eps paths
        f2[m0][9, : f2[m0].shape[1] // 2].zero_()   # shared half only
    elif kind == "dup_rows":
        # the sampler pads short subsequences by repeating the last sample
        # (multi_modal_dataset.py:105-106): duplicate neighbouring rows, zero distances
        for f in (f1, f2):
            for m in f:
                f[m][S * 3 + 3] = f[m][S * 3 + 2]
                f[m][S * 7 + 2] = f[m][S * 7 + 1]
                f[m][S * 7 + 3] = f[m][S * 7 + 1]
    return f1, f2


def build_inputs(case, dtype=torch.float32):
    S = case.get("seq_len", 4)
    if case["gen"] == "iid":
        f1, f2 = make_iid(case["seed"], case["mods"], case["B"], case["D"], dtype)
    else:
        f1, f2 = make_structured(case["seed"], case["mods"], case["B"], case["D"], S, dtype)
    if case.get("mutate"):
        f1, f2 = mutate_inputs(case["mutate"], f1, f2, S)
    return f1, f2


def reference_args(case):
    temperature = case.get("temperature", dict(DEFAULT_T))
    return types.SimpleNamespace(
        device="cpu", model=case["model"], tag=case.get("tag"),
        dataset_config={
            "modality_names": list(case["mods"]), "seq_len": case.get("seq_len", 4),
            "FOCAL": {"temperature": temperature, "inter_rank_margin": 1,
                      "shared_contrastive_loss_weight": 1, "private_contrastive_loss_weight": 1,
                      "orthogonal_loss_weight": 3, "rank_loss_weight": 5},
        })


def run_reference(case, dtype):
    from models.loss import FOCALLoss                   # the live reference
    from models.FOCALModules import split_features

    args = reference_args(case)
    ref = FOCALLoss(args)
    f1, f2 = build_inputs(case)                         # always the fp32 draw
    f1 = {m: v.to(dtype).requires_grad_(True) for m, v in f1.items()}
    f2 = {m: v.to(dtype).requires_grad_(True) for m, v in f2.items()}
    loss = ref(f1, f2)
    loss.backward()
    out = {"loss": loss.detach().numpy()}
    for m in case["mods"]:
        out[f"g1_{m}"] = f1[m].grad.numpy()
        out[f"g2_{m}"] = f2[m].grad.numpy()

    # un-weighted sub-loss sums, by driving the reference's own methods in its loop order
    with torch.no_grad():
        S = args.dataset_config["seq_len"]
        mods = case["mods"]
        r1 = {m: f1[m].reshape(-1, S, f1[m].shape[-1]) for m in mods}
        r2 = {m: f2[m].reshape(-1, S, f2[m].shape[-1]) for m in mods}
        s1, s2 = split_features(r1), split_features(r2)
        shared = 0.0
        for full, sp in ((r1, s1), (r2, s2)):
            for i, a in enumerate(mods):
                for c in mods[i + 1:]:
                    if case.get("tag") == "noPrivate":
                        shared += ref.forward_contrastive_loss(full[a], full[c])
                    else:
                        shared += ref.forward_contrastive_loss(sp[a]["shared"], sp[c]["shared"])
        private = sum(ref.forward_contrastive_loss(s1[m]["private"], s2[m]["private"]) for m in mods)
        temporal = sum(ref.forward_temporal_inter_ranking_loss(r[m]) for r in (r1, r2) for m in mods)
        orth = 0.0
        for sp in (s1, s2):
            for i, a in enumerate(mods):
                orth += ref.forward_orthogonality_loss(sp[a]["shared"], sp[a]["private"])
                for c in mods[i + 1:]:
                    orth += ref.forward_orthogonality_loss(sp[a]["private"], sp[c]["private"])
        out["parts"] = np.array([float(shared), float(private), float(orth), float(temporal)], dtype=np.float64)
    return out


def main():
    if not os.path.isdir(REF_SRC):
        raise SystemExit(f"{REF_SRC} not present: golden fixtures can only be regenerated in the build container")
    sys.path.insert(0, REF_SRC)
    torch.manual_seed(0)
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    outdir = os.path.join(ROOT, "tests", "golden")
    os.makedirs(outdir, exist_ok=True)
    only = set(sys.argv[1:])          # optional: regenerate just the named cases
    for case in CASES:
        if only and case["name"] not in only:
            continue
        f1, f2 = build_inputs(case)
        rec = {}
        r32 = run_reference(case, torch.float32)
        r64 = run_reference(case, torch.float64)
        rec["loss_f32"] = r32["loss"].astype(np.float32)
        rec["loss_f64"] = r64["loss"].astype(np.float64)
        rec["parts_f32"] = r32["parts"]
        rec["parts_f64"] = r64["parts"]
        for m in case["mods"]:
            rec[f"g1_{m}"] = r64[f"g1_{m}"].astype(np.float32)     # fp64 reference grads, stored as fp32
            rec[f"g2_{m}"] = r64[f"g2_{m}"].astype(np.float32)
            rec[f"g1f32_{m}_row0"] = r32[f"g1_{m}"][0].astype(np.float32)
        gn32 = np.sqrt(sum(float((r32[k].astype(np.float64) ** 2).sum()) for k in r32 if k.startswith("g")))
        gn64 = np.sqrt(sum(float((r64[k] ** 2).sum()) for k in r64 if k.startswith("g")))
        rec["gradnorm_f32"] = np.float64(gn32)
        rec["gradnorm_f64"] = np.float64(gn64)
        # inputs: stored for small cases, check-summed for all (RNG stream drift tripwire)
        small = case["B"] * case["D"] * len(case["mods"]) <= 128 * 128 * 2
        for m in case["mods"]:
            rec[f"chk1_{m}"] = np.array([f1[m].double().sum().item(), f1[m].double().abs().sum().item()])
            rec[f"chk2_{m}"] = np.array([f2[m].double().sum().item(), f2[m].double().abs().sum().item()])
            if small:
                rec[f"x1_{m}"] = f1[m].numpy()
                rec[f"x2_{m}"] = f2[m].numpy()
        rec["meta"] = np.array(repr({k: v for k, v in case.items()}))
        rec["torch_version"] = np.array(torch.__version__)
        np.savez_compressed(os.path.join(outdir, case["name"] + ".npz"), **rec)
        print(f"{case['name']:18s} loss32={float(rec['loss_f32']):.8f} loss64={float(rec['loss_f64']):.10f} "
              f"|g|32={gn32:.9f} parts={np.array2string(rec['parts_f64'], precision=6)}")


if __name__ == "__main__":
    main()
