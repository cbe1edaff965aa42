"""Helpers shared by the tests: load a golden fixture and rebuild its seeded inputs."""
import ast
import os

import numpy as np
import torch

from oracle.focal_oracle import FocalConfig
from oracle.make_golden import CASES, DEFAULT_T, build_inputs

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASE_BY_NAME = {c["name"]: c for c in CASES}
FINITE_CASES = [c["name"] for c in CASES if not c["name"].endswith("_nan")]


def load_case(name):
    case = CASE_BY_NAME[name]
    rec = np.load(os.path.join(GOLDEN_DIR, name + ".npz"), allow_pickle=False)
    assert ast.literal_eval(str(rec["meta"])) == case, "fixture was generated from a different case definition"
    f1, f2 = build_inputs(case)
    for m in case["mods"]:
        # RNG-stream tripwire: the regenerated inputs must be the ones the reference saw
        for f, key in ((f1, f"chk1_{m}"), (f2, f"chk2_{m}")):
            chk = np.array([f[m].double().sum().item(), f[m].double().abs().sum().item()])
            assert np.allclose(chk, rec[key], rtol=1e-12, atol=1e-9), f"seeded inputs drifted for {name}:{m}"
        if f"x1_{m}" in rec.files:
            assert np.array_equal(f1[m].numpy(), rec[f"x1_{m}"])
            assert np.array_equal(f2[m].numpy(), rec[f"x2_{m}"])
    return case, rec, f1, f2


def config_of(case) -> FocalConfig:
    t = case.get("temperature", DEFAULT_T)
    if isinstance(t, dict):
        t = t[case["model"]]
    return FocalConfig(modalities=list(case["mods"]), seq_len=case.get("seq_len", 4), temperature=float(t),
                       margin=1.0, w_shared=1.0, w_private=1.0, w_orth=3.0, w_rank=5.0,
                       no_private=(case.get("tag") == "noPrivate"))


def rel_err(a: torch.Tensor, b: torch.Tensor) -> float:
    a = a.double().flatten()
    b = b.double().flatten()
    return float((a - b).norm() / b.norm().clamp_min(1e-300))


def golden_grads(case, rec):
    g1 = {m: torch.from_numpy(rec[f"g1_{m}"]) for m in case["mods"]}
    g2 = {m: torch.from_numpy(rec[f"g2_{m}"]) for m in case["mods"]}
    return g1, g2
