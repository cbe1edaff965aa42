"""GPU: the reference's OWN training code running on the B200 kernels through the drop-in boundary.

`from models.loss import FOCALLoss` (reference src/train_utils/model_selection.py:11) resolves to
focal_b200/dropin/models/loss.py; everything else -- backbone (stock PyTorch DeepSense), augmenter, FOCAL wrapper,
`calc_pretrain_loss` (src/train_utils/loss_calc_utils.py:1-22), optimizer, `train.py` itself -- is the unmodified
reference as placed under oracle/_ref by `python -m oracle.build_ref` (it travels to the GPU box like a built .so).
"""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu


def _require():
    assert torch.cuda.is_available(), "GPU tests selected (-m gpu) but no CUDA device is visible"
    from oracle.build_ref import ref_available
    if not ref_available() and not os.path.isdir("/root/reference/src"):
        pytest.fail("oracle/_ref is missing: run `python -m oracle.build_ref` (or __graft_entry__.build()) where "
                    "/root/reference exists, before shipping the tree to the GPU box")


def test_calc_pretrain_loss_three_steps_match_the_reference_loss(tmp_path):
    """3 optimiser steps of the reference's pretrain loop body with the stock DeepSense backbone.  At every step the
    features the backbone handed to the loss are captured and the reference's own FOCALLoss is evaluated on them (same
    device): loss within 1e-4, d loss / d features within 2e-3 (fp32 mode; the batch is small)."""
    _require()
    import focal_b200
    from oracle.build_ref import import_reference_loss
    from tools import ref_harness as rh
    src = rh.make_run_dir(str(tmp_path), n_seq=2, samples_per_seq=4)
    B = 64
    o = rh.build_pretrain_objects(src, batch_size=B)
    try:
        assert isinstance(o.loss_func, focal_b200.FOCALLoss), type(o.loss_func)      # init_loss_func picked OUR class
        assert len(o.loss_func.state_dict()) == 0
        ref_loss_fn = import_reference_loss()(o.args).to(o.args.device)
        captured = {}
        inner = o.loss_func

        def recording_loss(f1, f2, *a, **k):
            for f in (f1, f2):
                for t in f.values():
                    t.retain_grad()
            captured["f"] = (f1, f2)
            return inner(f1, f2, *a, **k)

        o.default_model.train()
        losses = []
        for step in range(3):
            o.optimizer.zero_grad()
            loss = o.calc_pretrain_loss(o.args, o.default_model, o.augmenter, recording_loss,
                                        rh.synthetic_time_inputs(B, seed=step))
            loss.backward()
            o.optimizer.step()
            f1, f2 = captured["f"]
            r1 = {m: v.detach().clone().requires_grad_(True) for m, v in f1.items()}
            r2 = {m: v.detach().clone().requires_grad_(True) for m, v in f2.items()}
            ref = ref_loss_fn(r1, r2)
            ref.backward()
            torch.cuda.synchronize()
            assert abs(float(loss) - float(ref)) / abs(float(ref)) < 1e-4, (step, float(loss), float(ref))
            for fs, rs in ((f1, r1), (f2, r2)):
                for m in fs:
                    err = float((fs[m].grad - rs[m].grad).norm() / rs[m].grad.norm())
                    assert err < 2e-3, (step, m, err)
            losses.append(float(loss))
        assert all(map(lambda v: v == v, losses))
        assert inner.engine.graph_captures == 1 and inner.engine.graph_replays >= 2      # fresh activations every step
    finally:
        o.env.__exit__(None, None, None)


def test_reference_train_py_runs_unmodified_with_the_shadowed_loss(tmp_path):
    """`cd src; python train.py -model=DeepSense -dataset=MOD -learn_framework=FOCAL -batch_size=32` -- one epoch on a
    tiny synthetic MOD-format dataset, including the reference's KNN validation (eval under no_grad)."""
    _require()
    import focal_b200
    from focal_b200 import engine
    from tools import ref_harness as rh
    src = rh.make_run_dir(str(tmp_path), n_seq=8, samples_per_seq=8, epochs=1)
    before = dict(engine.CALLS)
    out = rh.run_train_literal(src, extra_argv=["-batch_size=32"], workers=None)
    assert out["loss_class"] is focal_b200.FOCALLoss
    assert os.path.normpath(out["loss_module"]).endswith(os.path.join("focal_b200", "dropin", "models", "loss.py"))
    ran = {k: engine.CALLS[k] - before.get(k, 0) for k in engine.CALLS}
    assert ran["grad"] >= 2 and ran["nograd"] >= 2, ran          # 2 training batches; val + test batches under no_grad
    logs = [os.path.join(r, f) for r, _, fs in os.walk(os.path.join(str(tmp_path), "weights")) for f in fs
            if f == "pretrain_log.txt"]
    assert logs, "the reference's pretrain log was not written"
    text = open(logs[0]).read()
    assert "Train contrastive loss" in text and "nan" not in text.lower()
