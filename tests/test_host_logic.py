"""CPU: host-side logic of the module / engine that does not need a device."""
import types

import pytest
import torch

from focal_b200 import FOCALLoss
from focal_b200.engine import FocalEngine, FocalHyper, shard_sequences


def make_args(mods=("seismic", "audio"), temp=None, model="DeepSense", tag=None, seq_len=4):
    temp = temp if temp is not None else {"SW_Transformer": 0.07, "DeepSense": 0.5}
    return types.SimpleNamespace(
        device="cpu", model=model, tag=tag,
        dataset_config={"modality_names": list(mods), "seq_len": seq_len,
                        "FOCAL": {"temperature": temp, "inter_rank_margin": 1, "ranking_margin": 0.0,
                                  "intra_rank_margin": 0.05, "shared_contrastive_loss_weight": 1,
                                  "private_contrastive_loss_weight": 1, "orthogonal_loss_weight": 3,
                                  "rank_loss_weight": 5}})


def test_module_mirrors_reference_constructor_contract():
    m = FOCALLoss(make_args())
    assert isinstance(m, torch.nn.Module)
    assert len(m.state_dict()) == 0 and list(m.parameters()) == [] and list(m.buffers()) == []
    assert m.temperature == 0.5 and m.modalities == ["seismic", "audio"]
    assert FOCALLoss(make_args(model="SW_Transformer")).temperature == pytest.approx(0.07)
    assert FOCALLoss(make_args(temp=0.2)).temperature == pytest.approx(0.2)          # scalar temperature
    m.to("cpu")                                                                      # init_loss_func calls .to(device)
    hp = m.engine.hp
    assert (hp.w_shared, hp.w_private, hp.w_orth, hp.w_rank, hp.margin) == (1.0, 1.0, 3.0, 5.0, 1.0)
    assert FOCALLoss(make_args(tag="noPrivate")).engine.hp.no_private is True
    assert len(FOCALLoss(make_args(mods=("a", "b", "c", "d", "e"))).modalities) == 5     # up to 8 modalities
    with pytest.raises(ValueError):
        FOCALLoss(make_args(mods=tuple("abcdefghi")))


def test_input_validation_raises_before_any_launch():
    m = FOCALLoss(make_args())
    ok = {k: torch.randn(32, 16) for k in ("seismic", "audio")}
    with pytest.raises(RuntimeError, match="CUDA"):
        m(ok, ok)                                                   # no CPU fallback
    bad_rows = {k: torch.randn(30, 16) for k in ("seismic", "audio")}
    with pytest.raises(ValueError):
        m(bad_rows, bad_rows)                                       # B % seq_len != 0 (reference: reshape raises)
    with pytest.raises(KeyError):
        m({"seismic": ok["seismic"]}, ok)
    half = {k: v.half() for k, v in ok.items()}
    with pytest.raises(RuntimeError, match="CUDA"):
        m(half, half)                                               # fp16 / bf16 activations are upcast, not refused
    ints = {k: v.long() for k, v in ok.items()}
    with pytest.raises(TypeError):
        m(ints, ints)
    mixed = {"seismic": torch.randn(32, 16), "audio": torch.randn(32, 8)}
    with pytest.raises(ValueError):
        m(mixed, mixed)


def test_shard_sequences():
    assert shard_sequences(2048, 8, 0) == (0, 256)
    assert shard_sequences(2048, 8, 7) == (1792, 2048)
    assert [shard_sequences(12, 3, r) for r in range(3)] == [(0, 4), (4, 8), (8, 12)]
    with pytest.raises(ValueError):
        shard_sequences(10, 4, 0)


def test_dropin_shadow_exports_the_same_class(repo_root):
    import importlib.util
    import os
    spec = importlib.util.spec_from_file_location("shadow_models_loss",
                                                  os.path.join(repo_root, "focal_b200", "dropin", "models", "loss.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    assert mod.FOCALLoss is FOCALLoss


def test_product_never_imports_the_oracle(repo_root):
    import os
    for root, _, files in os.walk(os.path.join(repo_root, "focal_b200")):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(root, fn)).read()
                assert "import oracle" not in text and "from oracle" not in text, os.path.join(root, fn)


def test_o_drain_turn_maps_eight_lanes_to_one_row():
    """The O-accumulator drain of the Gram kernels (gram_kernel.cuh: xchg_lane_reg_bit<0, 2>, <1, 3>, <2, 4>) restated in
    numpy: a 32 x 32 block held as v[lane = row][register = column] ends up with register 4 j + e of lane L holding
    row (L & 24) | j, column 4 (L & 7) + e -- so the 8 lanes L & 7 = 0..7 store the 128 bytes of one row."""
    import numpy as np
    rows, cols = np.meshgrid(np.arange(32), np.arange(32), indexing="ij")
    v = np.stack([rows, cols], axis=-1)                       # v[lane, reg] = (row, col) of the element held there

    def xchg(v, lb, rb):
        out = v.copy()
        for lane in range(32):
            hi = (lane >> lb) & 1
            partner = lane ^ (1 << lb)
            for a in range(32):
                if a & (1 << rb):
                    continue
                b = a | (1 << rb)
                send_of_partner = v[partner, a] if ((partner >> lb) & 1) else v[partner, b]
                if hi:
                    out[lane, a] = send_of_partner
                else:
                    out[lane, b] = send_of_partner
        return out

    for lb, rb in ((0, 2), (1, 3), (2, 4)):
        v = xchg(v, lb, rb)
    for lane in range(32):
        for j in range(8):
            for e in range(4):
                assert tuple(v[lane, 4 * j + e]) == ((lane & 24) | j, 4 * (lane & 7) + e)
    # every element exactly once
    assert len({tuple(x) for x in v.reshape(-1, 2)}) == 1024
