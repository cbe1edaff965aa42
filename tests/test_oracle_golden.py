"""CPU: the oracle (closed form + op-for-op port) against fixtures produced by the live reference."""
import math

import numpy as np
import pytest
import torch

from oracle import focal_oracle as fo
from oracle.focal_ref_port import focal_loss_port
from tests._golden import FINITE_CASES, config_of, golden_grads, load_case, rel_err

# Appendix B of SURVEY.md: values printed by the live reference in the survey container.
KAT = {
    "kat1_cfg1": (36.99359512, 0.731836291, [1.289398e-03, 1.328427e-03, 4.407059e-04]),
    "kat2": (39.87808228, 0.5046426782, [8.300106e-04, -1.300952e-03, -1.926150e-04]),
    "kat3_m3_t007": (75.98975372, 1.009898096, [-9.094857e-04, -6.682471e-04, 3.021116e-03]),
    "kat4_m4": (101.5352554, 1.499623677, [-2.419259e-03, -5.298733e-04, -1.654175e-02]),
    "kat5_noprivate": (34.03916168, 1.055361839, [-6.338627e-03, 4.010968e-03, 6.800976e-04]),
    "skat1": (16.0701313, 1.228713262, [3.649441e-03, 4.009286e-03, 3.300252e-03]),
    "skat2_m3_t007": (7.036208153, 1.432481432, [2.948113e-05, 7.815151e-04, 1.596018e-03]),
    "skat3": (20.16216469, 0.5764543234, [-2.291333e-05, -1.015147e-04, -1.739521e-05]),
}


@pytest.mark.parametrize("name", sorted(KAT))
def test_fixtures_match_survey_kat(name):
    case, rec, f1, f2 = load_case(name)
    loss, gnorm, g0 = KAT[name]
    assert float(rec["loss_f32"]) == pytest.approx(loss, rel=2e-7)
    assert float(rec["gradnorm_f32"]) == pytest.approx(gnorm, rel=2e-6)
    assert np.allclose(rec[f"g1f32_{case['mods'][0]}_row0"][:3], g0, rtol=2e-4, atol=1e-9)


@pytest.mark.parametrize("name", FINITE_CASES)
def test_closed_form_fp64_matches_reference(name):
    case, rec, f1, f2 = load_case(name)
    cfg = config_of(case)
    res = fo.focal_closed_form(f1, f2, cfg, dtype=torch.float64)
    assert float(res.loss) == pytest.approx(float(rec["loss_f64"]), rel=1e-9)
    parts = [float(res.parts[k]) for k in ("shared", "private", "orth", "temporal")]
    assert np.allclose(parts, rec["parts_f64"], rtol=1e-8, atol=1e-12)
    g1, g2 = golden_grads(case, rec)
    # fixtures store the fp64 reference gradients rounded to fp32 (6e-8 relative)
    tol = 2e-7
    if name == "edge_zero_row":
        tol = 1e-6        # 1/1e-8 amplification at the clamped zero-norm row
    for m in case["mods"]:
        assert rel_err(res.grads1[m], g1[m]) < tol, (name, m)
        assert rel_err(res.grads2[m], g2[m]) < tol, (name, m)


@pytest.mark.parametrize("name", ["kat1_cfg1", "kat4_m4", "kat5_noprivate", "skat1", "edge_odd_d", "edge_seq2",
                                  "edge_dup_rows"])
def test_closed_form_matches_autograd(name):
    case, rec, f1, f2 = load_case(name)
    cfg = config_of(case)
    a = fo.focal_closed_form(f1, f2, cfg, dtype=torch.float64)
    b = fo.focal_autograd(f1, f2, cfg, dtype=torch.float64)
    assert float(a.loss) == pytest.approx(float(b.loss), rel=1e-12)
    # duplicated rows: the Gram-form distance of identical rows is rounding noise, not exactly 0
    tol = 1e-8 if name == "edge_dup_rows" else 1e-11
    for m in case["mods"]:
        assert rel_err(a.grads1[m], b.grads1[m]) < tol
        assert rel_err(a.grads2[m], b.grads2[m]) < tol


@pytest.mark.parametrize("name", ["kat1_cfg1", "kat3_m3_t007", "skat1", "edge_odd_d", "edge_scalar_temp",
                                  "edge_ragged_b", "kat5_noprivate"])
def test_closed_form_fp32_within_north_star_tolerance(name):
    case, rec, f1, f2 = load_case(name)
    res = fo.focal_closed_form(f1, f2, config_of(case), dtype=torch.float32)
    assert abs(float(res.loss) - float(rec["loss_f64"])) / abs(float(rec["loss_f64"])) < 1e-5
    g1, g2 = golden_grads(case, rec)
    for m in case["mods"]:
        assert rel_err(res.grads1[m], g1[m]) < 1e-4
        assert rel_err(res.grads2[m], g2[m]) < 1e-4


@pytest.mark.parametrize("name", ["kat1_cfg1", "kat4_m4", "kat5_noprivate", "skat1", "edge_odd_d",
                                  "edge_scalar_temp", "edge_seq2"])
def test_op_for_op_port_matches_reference_fp32(name):
    """The timed CPU baseline is the reference's computation: same ATen ops, same fp32 results."""
    case, rec, f1, f2 = load_case(name)
    cfg = config_of(case)
    f1 = {m: v.clone().requires_grad_(True) for m, v in f1.items()}
    f2 = {m: v.clone().requires_grad_(True) for m, v in f2.items()}
    loss = focal_loss_port(f1, f2, cfg)
    loss.backward()
    assert float(loss.detach()) == pytest.approx(float(rec["loss_f32"]), rel=1e-6)
    gn = math.sqrt(sum(float((f[m].grad.double() ** 2).sum()) for f in (f1, f2) for m in case["mods"]))
    assert gn == pytest.approx(float(rec["gradnorm_f32"]), rel=1e-5)
    g1, g2 = golden_grads(case, rec)
    for m in case["mods"]:
        assert rel_err(f1[m].grad, g1[m]) < 2e-5


@pytest.mark.parametrize("name", ["edge_b1_nan", "edge_seq1_nan"])
def test_degenerate_batches_are_nan_like_the_reference(name):
    case, rec, f1, f2 = load_case(name)
    assert math.isnan(float(rec["loss_f32"]))
    res = fo.focal_closed_form(f1, f2, config_of(case), dtype=torch.float64, need_grad=False)
    assert math.isnan(float(res.loss))
    # the InfoNCE / orth parts stay finite (Appendix E)
    assert np.allclose([float(res.parts[k]) for k in ("shared", "private", "orth")], rec["parts_f64"][:3],
                       rtol=1e-8, atol=1e-12)


def test_batch_not_multiple_of_seq_len_raises():
    f1, f2 = fo.make_iid(0, ["a", "b"], 30, 16)
    with pytest.raises(ValueError):
        fo.focal_closed_form(f1, f2, fo.FocalConfig(modalities=["a", "b"], seq_len=4))


def test_index_contract():
    """Integer contract: positives, sequence membership, problem lists."""
    b, S = 5, 4
    N = 2 * b
    for k in range(N):
        p = fo.positive_index(k, b)
        assert p != k and fo.positive_index(p, b) == k and (p % b) == (k % b)
    assert [fo.sequence_of_row(i, S) for i in (0, 3, 4, 19)] == [(0, 0), (0, 3), (1, 0), (4, 3)]
    for M in (2, 3, 4):
        probs = fo.nce_problem_list(M)
        assert len(probs) == M * M
        assert sum(p[0] == "private" for p in probs) == M
        assert len(fo.orth_pair_list(M)) == 2 * (M + M * (M - 1) // 2)
    # reference order for M=3, view 0: (0,1) (0,2) (1,2)
    assert [(p[2], p[4]) for p in fo.nce_problem_list(3)[:3]] == [(0, 1), (0, 2), (1, 2)]


def test_row_shards_sum_to_global():
    """Multi-GPU semantics: per-shard terms add up to the single-device result."""
    case, rec, f1, f2 = load_case("skat1")
    cfg = config_of(case)
    full = fo.focal_closed_form(f1, f2, cfg, dtype=torch.float64)
    b = case["B"] // cfg.seq_len
    cuts = [0, 8, 20, b]
    loss = 0.0
    g = {m: torch.zeros_like(full.grads1[m]) for m in case["mods"]}
    for lo, hi in zip(cuts[:-1], cuts[1:]):
        part = fo.focal_closed_form(f1, f2, cfg, dtype=torch.float64, seq_rows=(lo, hi))
        loss += float(part.loss)
        for m in case["mods"]:
            rows = slice(lo * cfg.seq_len, hi * cfg.seq_len)
            assert float(part.grads1[m][: lo * cfg.seq_len].abs().sum()) == 0.0
            g[m][rows] = part.grads1[m][rows]
    assert loss == pytest.approx(float(full.loss), rel=1e-12)
    for m in case["mods"]:
        assert rel_err(g[m], full.grads1[m]) < 1e-12
