"""GPU parity: the CUDA path (through the module API -> ctypes -> C ABI -> sm_100a kernels) against the oracle
and the golden fixtures of the live reference.

Tolerances are north_star's: loss 1e-4 relative; gradients (bf16 tile mode) 1e-2 relative per tensor, norm-wise.
Integer results (active hinge counts' index sets on well-separated inputs, positive indices) are exact.
"""
import math
import types

import numpy as np
import pytest
import torch

from oracle import focal_oracle as fo
from tests._golden import CASE_BY_NAME, FINITE_CASES, config_of, golden_grads, load_case, rel_err

pytestmark = pytest.mark.gpu

LOSS_RTOL = 1e-4            # north_star: both precision modes, every fixture
GRAD_RTOL = {"fp32": 2e-3, "bf16": 1e-2}
# |m_II - m_IJ + margin| below this may flip its active flag relative to the fp64 oracle: bf16 tiles move block means by
# ~1e-3; the fp32 mode (split tiles) by ~1e-6, like the fp32 reference itself
HINGE_TOL = {"fp32": 5e-5, "bf16": 4e-3}
MAX_HINGE_FLIPS = 8         # per tensor; more than a handful means the distances are wrong, not a kink effect


def _require_cuda():
    assert torch.cuda.is_available(), "GPU tests selected (-m gpu) but no CUDA device is visible"


def make_args(cfg: fo.FocalConfig, model="DeepSense", scalar_temp=True, precision="auto"):
    temp = cfg.temperature if scalar_temp else {model: cfg.temperature}
    return types.SimpleNamespace(
        device="cuda", model=model, tag="noPrivate" if cfg.no_private else None, focal_precision=precision,
        dataset_config={"modality_names": list(cfg.modalities), "seq_len": cfg.seq_len,
                        "FOCAL": {"temperature": temp, "inter_rank_margin": cfg.margin,
                                  "shared_contrastive_loss_weight": cfg.w_shared,
                                  "private_contrastive_loss_weight": cfg.w_private,
                                  "orthogonal_loss_weight": cfg.w_orth, "rank_loss_weight": cfg.w_rank}})


def resolved_mode(precision: str, B: int, D: int) -> str:
    from focal_b200 import _cabi
    from focal_b200.engine import FocalHyper, resolve_precision
    hp = FocalHyper(("a",), 4, 0.5, 1.0, 1.0, 1.0, 3.0, 5.0, False, 7, precision)
    return "fp32" if resolve_precision(hp, B, D) == _cabi.FOCAL_PREC_FP32 else "bf16"


def flip_correction(ref: fo.FocalResult, x64: torch.Tensor, t: int, S: int, cnt_ours: torch.Tensor, mode: str, cfg):
    """Hinge flags that flipped across the kink (the loss is continuous there, the gradient is
    not): found from the per-sequence active counts the kernel publishes, attributed to the borderline pairs of that
    sequence, and turned into the EXACT gradient change they cause (the gradient is linear in the active set).
    Returns (number of flips, gradient correction [B, D]) -- or raises when a count difference has no borderline pair
    to explain it."""
    ax = ref.aux["temporal"][t]
    m, act = ax["m"].cpu(), ax["active"].cpu()
    b = m.shape[0]
    h = ax["mII"].cpu()[:, None] - m + cfg.margin
    off = ~torch.eye(b, dtype=torch.bool)
    d = cnt_ours.long().cpu() - act.sum(dim=1)
    delta = torch.zeros(b, b, dtype=torch.float64)
    for I in d.nonzero().flatten().tolist():
        need = abs(int(d[I]))
        # ours has MORE active flags than the oracle -> oracle-inactive borderline pairs became active, and vice versa
        cand = (h[I].abs() < HINGE_TOL[mode]) & off[I] & (act[I] == (d[I] < 0))
        idx = cand.nonzero().flatten()
        assert idx.numel() >= need, (f"tensor {t} sequence {I}: active-hinge count differs by {int(d[I])} but only "
                                     f"{idx.numel()} pairs lie within {HINGE_TOL[mode]} of the kink")
        idx = idx[h[I, idx].abs().argsort()[:need]]
        delta[I, idx] = 1.0 if d[I] > 0 else -1.0
    n = int(delta.abs().sum())
    if n == 0:
        return 0, None
    return n, cfg.w_rank * fo.temporal_gradient_of_active_set(x64.cpu().double(), S, delta)


def assert_grads_close(got, want, ref, x, t, S, tol, what, cnt_ours, mode, cfg):
    """Norm-wise per-tensor comparison.  When it fails, the only accepted explanation is a bounded number of hinge flags
    that flipped across the kink; their exact gradient contribution is added to the reference and the comparison must
    then hold on ALL rows (nothing is masked)."""
    want = want.to(got.device)
    err = rel_err(got, want)
    if err < tol:
        return 0
    n, corr = flip_correction(ref, x, t, S, cnt_ours, mode, cfg)
    assert 0 < n <= MAX_HINGE_FLIPS, (what, err, f"{n} hinge flips: not a kink effect")
    err2 = rel_err(got, want + corr.to(got.device, got.dtype if got.dtype == torch.float64 else torch.float64))
    assert err2 < tol, (what, f"error {err:.2e}; {err2:.2e} after accounting for {n} flipped hinge flag(s)")
    return n


def run_module(f1, f2, cfg, need_grad=True, precision="auto"):
    """Runs the module API once.  Returns (module, loss, grads1, grads2, cnt): cnt = active hinge count per
    (tensor, sequence) as published by the temporal kernel (None when the temporal term is degenerate)."""
    from focal_b200 import FOCALLoss
    mod = FOCALLoss(make_args(cfg, precision=precision)).to("cuda")
    g1 = {m: v.cuda().requires_grad_(need_grad) for m, v in f1.items()}
    g2 = {m: v.cuda().requires_grad_(need_grad) for m, v in f2.items()}
    x0 = next(iter(g1.values()))
    B, D = x0.shape
    be, hp = mod.engine.backend, mod.engine.hp
    try:
        _, ws, info, _, _ = be.plan(hp, B, D, need_grad, (0, B // cfg.seq_len), x0.device)
        ws.zero_()                 # stream-K piece copies that a launch does not write must read as zero below
    except ValueError:
        ws = info = None           # unsupported shape: let the module raise
    if need_grad:
        loss = mod(g1, g2)
        loss.backward()
    else:
        with torch.no_grad():
            loss = mod(g1, g2)
    torch.cuda.synchronize()
    cnt = None
    if ws is not None and need_grad:
        nT = 2 * len(cfg.modalities)
        cnt = sum(be._view(ws, info.cnt_off + k * info.cnt_piece_stride, info.cnt_bytes, torch.int32,
                           (nT, info.bpad)).cpu() for k in range(info.n_pieces_tmp))[:, :info.b]
    return mod, loss.detach().cpu(), g1, g2, cnt


def check_all_grads(name, mods, g1, g2, r1, r2, ref, f1, f2, cfg, cnt, mode):
    M = len(mods)
    flips = 0
    for i, m in enumerate(mods):
        tol = GRAD_RTOL[mode]
        flips += assert_grads_close(g1[m].grad.cpu(), r1[m], ref, f1[m], i, cfg.seq_len, tol, (name, m, 1),
                                    None if cnt is None else cnt[i], mode, cfg)
        flips += assert_grads_close(g2[m].grad.cpu(), r2[m], ref, f2[m], M + i, cfg.seq_len, tol, (name, m, 2),
                                    None if cnt is None else cnt[M + i], mode, cfg)
    return flips


@pytest.mark.parametrize("precision", ["auto", "fp32", "bf16"])
@pytest.mark.parametrize("name", FINITE_CASES)
def test_golden_cases(name, precision):
    """Every fixture of the live reference: loss, the four sub-losses, all gradients -- in the default ("auto") mode, in
    the fp32 mode (gradients 2e-3) and in the bf16 mode (1e-2).  The loss bar is 1e-4 everywhere."""
    _require_cuda()
    case, rec, f1, f2 = load_case(name)
    cfg = config_of(case)
    if precision == "fp32" and case["D"] > 256:
        pytest.skip("fp32 mode (split tiles) stops at D = 256; auto runs these in bf16 mode")
    if precision == "bf16" and case["B"] * case["D"] < 128 * 96:
        # bf16 tiles evaluate the loss at features rounded to 8 significant bits; the induced error (grad . eta) is
        # averaged down by sqrt(B D), which these tiny edge-case batches do not have -- the default mode runs them
        # (and everything up to 2048 rows) with split tiles
        pytest.skip("bf16 mode is not selected for batches this small (auto -> fp32 mode)")
    mode = resolved_mode(precision, case["B"], case["D"])
    mod, loss, g1, g2, cnt = run_module(f1, f2, cfg, precision=precision)
    ref_loss = float(rec["loss_f64"])
    assert abs(float(loss) - ref_loss) / abs(ref_loss) < LOSS_RTOL, (float(loss), ref_loss)
    parts = mod.last_parts.cpu().double().numpy()[1:]
    assert np.allclose(parts, rec["parts_f64"], rtol=2 * LOSS_RTOL, atol=2e-5), (parts, rec["parts_f64"])
    r1, r2 = golden_grads(case, rec)
    ref = fo.focal_closed_form(f1, f2, cfg, dtype=torch.float64)      # only for the hinge-kink bookkeeping
    check_all_grads(name, case["mods"], g1, g2, r1, r2, ref, f1, f2, cfg, cnt, mode)


@pytest.mark.parametrize("name", ["edge_b1_nan", "edge_seq1_nan"])
def test_degenerate_batches(name):
    """b == 1 or S == 1: NaN loss like the reference, finite InfoNCE / orthogonality parts and gradients."""
    _require_cuda()
    case, rec, f1, f2 = load_case(name)
    mod, loss, g1, g2, _ = run_module(f1, f2, config_of(case))
    mode = resolved_mode("auto", case["B"], case["D"])
    assert math.isnan(float(loss))
    parts = mod.last_parts.cpu().double().numpy()[1:]
    assert np.allclose(parts[:3], rec["parts_f64"][:3], rtol=2e-4, atol=2e-5)
    r1, r2 = golden_grads(case, rec)
    for m in case["mods"]:
        assert torch.isfinite(g1[m].grad).all()
        assert rel_err(g1[m].grad.cpu(), r1[m]) < GRAD_RTOL[mode]


def oracle_check(name, f1, f2, cfg, mods, precision):
    """Module API vs the fp64 closed-form oracle (evaluated on the GPU): loss 1e-4, gradients at the mode's tolerance."""
    B, D = next(iter(f1.values())).shape
    mode = resolved_mode(precision, B, D)
    mod, loss, g1, g2, cnt = run_module(f1, f2, cfg, precision=precision)
    ref = fo.focal_closed_form({m: v.cuda() for m, v in f1.items()}, {m: v.cuda() for m, v in f2.items()}, cfg,
                               dtype=torch.float64)
    assert abs(float(loss) - float(ref.loss)) / abs(float(ref.loss)) < LOSS_RTOL, (float(loss), float(ref.loss))
    r1 = {m: v.cpu() for m, v in ref.grads1.items()}
    r2 = {m: v.cpu() for m, v in ref.grads2.items()}
    return check_all_grads(name, mods, g1, g2, r1, r2, ref, f1, f2, cfg, cnt, mode)


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
@pytest.mark.parametrize("gen,B,D,mods,T,seed", [
    ("iid", 1024, 256, ["seismic", "audio"], 0.5, 0),            # cfg 2 shape
    ("structured", 1024, 256, ["seismic", "audio"], 0.5, 1),
    ("structured", 2048, 256, ["acc", "gyr", "mag"], 0.07, 2),     # cfg 3 shape (3 modalities, T = 0.07)
    ("structured", 4096, 256, ["acc", "gyr", "mag"], 0.07, 9),     # cfg 3 at its full batch
    ("iid", 1536, 128, ["seismic", "audio"], 0.5, 3),            # b = 384: three row tiles, K blocks = 1 / 2
    ("structured", 4 * 333, 192, ["a", "b"], 0.2, 4),             # ragged: b = 333, D = 192 (3 K blocks, BN = 64)
    ("iid", 2048, 64, ["m0", "m1", "m2", "m3"], 0.5, 5),          # 4 modalities
    ("structured", 1024, 512, ["seismic", "audio"], 0.5, 6),      # cfg 5 width (256 + 256): wide temporal mode, 8 K blocks
    ("iid", 4 * 200, 512, ["a", "b", "c", "d"], 0.5, 7),          # ... with 4 modalities and a ragged batch
    ("iid", 512, 320, ["a", "b"], 0.5, 8),                        # 256 < D < 512: zero-padded to 8 K blocks
])
def test_against_fp64_oracle(gen, B, D, mods, T, seed, precision):
    """Sizes the reference cannot hold in memory comfortably: compare with the fp64 closed-form oracle (on the GPU)."""
    _require_cuda()
    if precision == "fp32" and D > 256:
        pytest.skip("fp32 mode (split tiles) stops at D = 256")
    cfg = fo.FocalConfig(modalities=mods, seq_len=4, temperature=T)
    f1, f2 = (fo.make_iid(seed, mods, B, D) if gen == "iid" else fo.make_structured(seed, mods, B, D, 4))
    oracle_check((gen, B, D), f1, f2, cfg, mods, precision)


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
@pytest.mark.parametrize("S,B,D,tag", [
    (2, 256, 128, None),          # two rows per sequence (SEQ = 2 epilogue, fused intra means)
    (8, 512, 128, None),          # S = 8: generic finalize + separate intra-sequence kernel
    (16, 512, 64, None),
    (32, 1024, 96, None),         # a sequence spans a whole warp; D/2 = 48 takes the generic row kernels
    (4, 512, 256, "noPrivate"),   # shared InfoNCE on full-width rows (4 K blocks) + private on halves (2 K blocks)
    (4, 384, 128, "noPrivate"),
])
def test_sequence_lengths_and_noprivate(S, B, D, tag, precision):
    _require_cuda()
    if precision == "fp32" and tag == "noPrivate" and D > 128:
        pytest.skip("fp32 mode: full-width InfoNCE operands (noPrivate) stop at D = 128")
    mods = ["seismic", "audio"]
    cfg = fo.FocalConfig(modalities=mods, seq_len=S, temperature=0.5, no_private=(tag == "noPrivate"))
    f1, f2 = fo.make_structured(11 + S, mods, B, D, S)
    oracle_check((S, B, D, tag), f1, f2, cfg, mods, precision)


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_headline_size_against_fp64_oracle(precision):
    """BASELINE.json's metric configuration: B = 8192, M = 2, S = 4, D = 256, T = 0.5 -- both precision modes."""
    _require_cuda()
    mods = ["seismic", "audio"]
    cfg = fo.FocalConfig(modalities=mods, seq_len=4, temperature=0.5)
    f1, f2 = fo.make_structured(0, mods, 8192, 256, 4)
    oracle_check("headline", f1, f2, cfg, mods, precision)


@pytest.mark.parametrize("mode", ["fp32", "bf16"])
def test_intermediates_rowsum_and_hinge_counts(mode):
    """Index-exactness: row sums exclude exactly j == k; active hinge sets match the oracle on separated inputs."""
    _require_cuda()
    import ctypes as C
    from focal_b200 import _cabi
    from focal_b200.engine import CudaBackend, FocalHyper
    mods = ["seismic", "audio"]
    B, D, S = 640, 128, 4
    f1, f2 = fo.make_structured(3, mods, B, D, S)
    cfg = fo.FocalConfig(modalities=mods, seq_len=S, temperature=0.5)
    hp = FocalHyper(tuple(mods), S, 0.5, 1.0, 1.0, 1.0, 3.0, 5.0, False, 7, mode)
    be = CudaBackend()
    feats = [f1[m].cuda() for m in mods] + [f2[m].cuda() for m in mods]
    b = B // S
    loss5, grads = be.run(hp, feats, (0, b), True, None)
    torch.cuda.synchronize()
    ws, info = be.workspace(be._cfg(hp, B, D, True, (0, b)), feats[0].device)
    rs = be._view(ws, info.rowsum_off, info.rowsum_bytes, torch.float32, (info.n_problems, S, 2, info.bpad)).cpu()
    ref = fo.focal_closed_form(f1, f2, cfg, dtype=torch.float64)
    for q in range(info.n_problems):
        want = ref.aux["nce_rowsum"][q] * math.exp(1.0 / 0.5)          # oracle sums exp(s - 1/T)
        got = torch.cat((rs[q, :, 0, :b], rs[q, :, 1, :b]), dim=1).double()
        assert torch.allclose(got, want, rtol=3e-3 if mode == "bf16" else 2e-5), (q, float((got / want - 1).abs().max()))
    # one copy of the counts per stream-K piece of a row block (unused copies hold stale data -> only sum pieces that
    # exist: a piece that does not exist for a block was never written, so zero the workspace first and re-run)
    ws.zero_()
    be.run(hp, feats, (0, b), True, None)
    torch.cuda.synchronize()
    cnt = sum(be._view(ws, info.cnt_off + k * info.cnt_piece_stride, info.cnt_bytes, torch.int32,
                       (2 * len(mods), info.bpad)).cpu() for k in range(info.n_pieces_tmp))
    for t in range(2 * len(mods)):
        act = ref.aux["temporal"][t]["active"]
        m = ref.aux["temporal"][t]["m"]
        h = ref.aux["temporal"][t]["mII"][:, None] - m + 1.0
        want = act.sum(dim=1)
        got = cnt[t, :b].long()
        # pairs whose hinge is within bf16-tile noise of the kink may flip; all others must agree exactly
        border = ((h.abs() < HINGE_TOL[mode]) & ~torch.eye(b, dtype=torch.bool)).sum(dim=1)
        assert ((got - want).abs() <= border).all(), (t, int((got - want).abs().max()))


def test_forward_only_matches_and_skips_gradients():
    _require_cuda()
    case, rec, f1, f2 = load_case("skat1")
    cfg = config_of(case)
    _, loss_ng, g1, _, _ = run_module(f1, f2, cfg, need_grad=False)
    _, loss_g, _, _, _ = run_module(f1, f2, cfg, need_grad=True)
    assert float(loss_ng) == pytest.approx(float(loss_g), rel=1e-6)
    assert all(v.grad is None for v in g1.values())


def test_backward_scales_with_upstream_gradient_and_partial_requires_grad():
    _require_cuda()
    from focal_b200 import FOCALLoss
    case, rec, f1, f2 = load_case("kat1_cfg1")
    cfg = config_of(case)
    mod = FOCALLoss(make_args(cfg, scalar_temp=False)).to("cuda")
    assert len(mod.state_dict()) == 0 and sum(p.numel() for p in mod.parameters()) == 0
    a = {m: v.cuda().requires_grad_(True) for m, v in f1.items()}
    b2 = {m: v.cuda() for m, v in f2.items()}                      # view 2 does not require grad
    (mod(a, b2) * 3.0).backward()
    r1, _ = golden_grads(case, rec)
    for m in case["mods"]:
        assert rel_err(a[m].grad.cpu() / 3.0, r1[m]) < GRAD_RTOL[resolved_mode("auto", case["B"], case["D"])]


def test_determinism_bitwise():
    _require_cuda()
    case, rec, f1, f2 = load_case("skat3")
    cfg = config_of(case)
    _, l1, a1, _, _ = run_module(f1, f2, cfg)
    _, l2, a2, _, _ = run_module(f1, f2, cfg)
    assert float(l1) == float(l2)
    for m in case["mods"]:
        assert torch.equal(a1[m].grad, a2[m].grad)


def test_error_behaviour():
    _require_cuda()
    from focal_b200 import FOCALLoss
    cfg = fo.FocalConfig(modalities=["a", "b"], seq_len=4)
    mod = FOCALLoss(make_args(cfg))
    x = {m: torch.randn(30, 16, device="cuda") for m in ("a", "b")}
    with pytest.raises(ValueError):                       # B % S != 0: the reference's reshape raises too
        mod(x, x)
    y = {m: torch.randn(32, 16) for m in ("a", "b")}      # CPU tensors: no fallback
    with pytest.raises(RuntimeError):
        mod(y, y)
    z = {m: torch.randn(32, 600, device="cuda") for m in ("a", "b")}
    with pytest.raises(ValueError):                       # D > 512 not supported by the tile configurations
        mod(z, z)


def test_row_shards_on_one_gpu_sum_to_global():
    """The multi-GPU row sharding, exercised on one device: shard results add up to the unsharded result."""
    _require_cuda()
    from focal_b200.engine import CudaBackend, FocalHyper
    mods = ["seismic", "audio"]
    B, D, S = 1024, 128, 4
    f1, f2 = fo.make_structured(5, mods, B, D, S)
    hp = FocalHyper(tuple(mods), S, 0.5, 1.0, 1.0, 1.0, 3.0, 5.0)
    be = CudaBackend()
    feats = [f1[m].cuda() for m in mods] + [f2[m].cuda() for m in mods]
    b = B // S
    full5, fullg = be.run(hp, feats, (0, b), True, None)
    full5, fullg = full5.clone(), [g.clone() for g in fullg]
    acc5 = torch.zeros_like(full5)
    cuts = [0, 64, 128, 192, b]                            # 4 "ranks" (64 sequences = 256 rows each)
    sums = {}

    for lo, hi in zip(cuts[:-1], cuts[1:]):
        def exch(rs, lo=lo, hi=hi):
            sums[(lo, hi)] = rs[..., lo:hi].clone()
        be.run(hp, feats, (lo, hi), True, exch)
    allrs = None
    for lo, hi in zip(cuts[:-1], cuts[1:]):
        def exch(rs, lo=lo, hi=hi):
            for (a, c), v in sums.items():
                rs[..., a:c] = v
        l5, g = be.run(hp, feats, (lo, hi), True, exch)
        acc5 += l5
        for t in range(len(feats)):
            rows = slice(lo * S, hi * S)
            assert g[t].shape == ((hi - lo) * S, D)            # only the owned rows are returned
            assert rel_err(g[t], fullg[t][rows]) < 1e-5
    assert torch.allclose(acc5, full5, rtol=1e-5)


def test_cuda_graph_replay_matches_eager():
    """Same input buffers seen again -> the step is captured into a CUDA graph and replayed; results are identical."""
    _require_cuda()
    from focal_b200.engine import FocalEngine, FocalHyper
    mods = ["seismic", "audio"]
    B, D, S = 512, 128, 4
    f1, f2 = fo.make_structured(9, mods, B, D, S)
    hp = FocalHyper(tuple(mods), S, 0.5, 1.0, 1.0, 1.0, 3.0, 5.0)
    x1 = {m: v.cuda() for m, v in f1.items()}
    x2 = {m: v.cuda() for m, v in f2.items()}
    eager = FocalEngine(hp, use_cuda_graph=False)
    l_ref, g_ref = eager.loss_and_grads(x1, x2, True)
    l_ref, g_ref = l_ref.clone(), [g.clone() for g in g_ref]
    eng = FocalEngine(hp, use_cuda_graph=True)
    for it in range(4):
        l5, g = eng.loss_and_grads(x1, x2, True)
        torch.cuda.synchronize()
        assert torch.equal(l5, l_ref), it
        for a, b2 in zip(g, g_ref):
            assert torch.equal(a, b2), it
    assert eng.graph_replays >= 2
    # new data in the same buffers: the replay must pick it up
    for m in mods:
        x1[m].mul_(1.25)
    l5, g = eng.loss_and_grads(x1, x2, True)
    l_e, g_e = eager.loss_and_grads(x1, x2, True)
    torch.cuda.synchronize()
    assert torch.equal(l5, l_e) and all(torch.equal(a, b2) for a, b2 in zip(g, g_e))


def test_cuda_graph_serves_fresh_addresses_and_never_aliases_outputs():
    """A training loop hands the loss freshly allocated activations every step.  The captured launch sequence reads
    the caller's pointers from a table in the workspace (focal_b200_set_ptrs), so it replays for ANY address, its
    outputs are fresh tensors, and a second forward before the first backward cannot overwrite the first one's saved
    gradients (VERDICT r1 weak #5 / ADVICE r1)."""
    _require_cuda()
    from focal_b200 import FOCALLoss
    from focal_b200.engine import FocalEngine, FocalHyper
    mods = ["seismic", "audio"]
    B, D, S = 512, 128, 4
    cfg = fo.FocalConfig(modalities=mods, seq_len=S, temperature=0.5)
    hp = FocalHyper(tuple(mods), S, 0.5, 1.0, 1.0, 1.0, 3.0, 5.0)
    eager = FocalEngine(hp, use_cuda_graph=False)
    eng = FocalEngine(hp, use_cuda_graph=True)
    keep = []                                              # hold on to every input: each step sees new addresses
    for step in range(6):
        f1, f2 = fo.make_structured(20 + step, mods, B, D, S)
        x1 = {m: v.cuda() for m, v in f1.items()}
        x2 = {m: v.cuda() for m, v in f2.items()}
        keep.append((x1, x2))
        l5, g = eng.loss_and_grads(x1, x2, True)
        le, ge = eager.loss_and_grads(x1, x2, True)
        torch.cuda.synchronize()
        assert torch.equal(l5, le), step
        assert all(torch.equal(a, b2) for a, b2 in zip(g, ge)), step
    assert len({x1[mods[0]].data_ptr() for x1, _ in keep}) == 6
    assert eng.graph_captures == 1 and eng.graph_replays == 5
    # forward twice on the SAME buffers (refilled in place), then backward of the first loss
    mod = FOCALLoss(make_args(cfg)).to("cuda")
    a1 = {m: v.clone().requires_grad_(True) for m, v in keep[0][0].items()}
    a2 = {m: v.clone().requires_grad_(True) for m, v in keep[0][1].items()}
    for _ in range(3):                                     # warm the module's engine: eager, capture, replay
        mod(a1, a2)
    loss_a = mod(a1, a2)
    want = [g.clone() for g in eager.loss_and_grads(a1, a2, True)[1]]
    with torch.no_grad():
        for m in mods:
            a1[m].mul_(0.5)                                # new data in the same buffers
    loss_b = mod(a1, a2)                                   # replays the same graph
    loss_a.backward()                                      # must still see the gradients of the FIRST forward
    torch.cuda.synchronize()
    got = [a1[m].grad for m in mods] + [a2[m].grad for m in mods]
    assert all(torch.equal(a, b2) for a, b2 in zip(got, want))
    assert float(loss_a) != float(loss_b)
    assert mod.engine.graph_replays >= 3


def test_second_device_and_non_current_device():
    """Features on cuda:1 while cuda:0 is current (no device guard, per-process attribute cache)."""
    _require_cuda()
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    case, rec, f1, f2 = load_case("skat1")
    cfg = config_of(case)
    from focal_b200 import FOCALLoss
    mod = FOCALLoss(make_args(cfg))
    outs = []
    for dev in ("cuda:0", "cuda:1"):
        g1 = {m: v.to(dev).requires_grad_(True) for m, v in f1.items()}
        g2 = {m: v.to(dev).requires_grad_(True) for m, v in f2.items()}
        torch.cuda.set_device(0)
        loss = mod(g1, g2)
        loss.backward()
        torch.cuda.synchronize(dev)
        outs.append((float(loss), g1[case["mods"][0]].grad.cpu()))
    assert outs[0][0] == outs[1][0] and torch.equal(outs[0][1], outs[1][1])


def test_row_blocked_inputs_match_contiguous():
    """The layout an all-gather of per-rank [2M, B/R, D] buffers produces is read in place (no re-pack copy)."""
    _require_cuda()
    from focal_b200.engine import CudaBackend, FocalHyper
    mods = ["seismic", "audio"]
    B, D, S, R = 1024, 128, 4, 4
    f1, f2 = fo.make_structured(6, mods, B, D, S)
    hp = FocalHyper(tuple(mods), S, 0.5, 1.0, 1.0, 1.0, 3.0, 5.0)
    be = CudaBackend()
    feats = [f1[m].cuda() for m in mods] + [f2[m].cuda() for m in mods]
    nT, Bl = len(feats), B // R
    l_ref, g_ref = be.run(hp, feats, (0, B // S), True, None)
    l_ref, g_ref = l_ref.clone(), [g.clone() for g in g_ref]
    blocked = torch.stack([torch.stack([f[r * Bl:(r + 1) * Bl] for f in feats]) for r in range(R)])   # [R, 2M, Bl, D]
    heads = [blocked[0, t] for t in range(nT)]
    l5, g = be.run(hp, heads, (0, B // S), True, None, blocked=(B, Bl, nT * Bl * D))
    torch.cuda.synchronize()
    assert torch.equal(l5, l_ref)
    for a, b2 in zip(g, g_ref):
        assert torch.equal(a, b2)
