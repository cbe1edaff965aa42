"""CPU, build container only: the namespace-package shadowing really makes the UNMODIFIED reference pick up our loss.

/root/reference exists only in the build container (never on the GPU box), so this test is skipped elsewhere.  It runs
in a subprocess so that the path manipulation and the import stubs (matplotlib / timm are not installed here) stay out
of the test session.  What it proves: with ``focal_b200/dropin`` ahead of ``<reference>/src`` on sys.path, the
reference's own ``train_utils.model_selection.init_loss_func(args)`` (model_selection.py:47-59) returns an instance of
``focal_b200.loss.FOCALLoss`` while ``models.DeepSense`` / ``models.FOCALModules`` still come from the reference.
"""
import os
import subprocess
import sys
import textwrap

import pytest

REF = "/root/reference/src"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SCRIPT = textwrap.dedent('''
    import sys, types
    import torch
    root, ref = sys.argv[1], sys.argv[2]
    sys.path[:0] = [root, root + "/focal_b200/dropin", ref]
    # import-time stubs for packages the reference imports but this image lacks (SURVEY.md Appendix D)
    mpl = types.ModuleType("matplotlib"); plt = types.ModuleType("matplotlib.pyplot"); plt.axis = lambda *a, **k: None
    mpl.pyplot = plt; sys.modules["matplotlib"] = mpl; sys.modules["matplotlib.pyplot"] = plt
    timm = types.ModuleType("timm"); tm = types.ModuleType("timm.models"); tl = types.ModuleType("timm.models.layers")
    tl.trunc_normal_ = torch.nn.init.trunc_normal_
    tl.DropPath = torch.nn.Identity
    tl.to_2tuple = lambda x: (x, x)
    timm.models = tm; tm.layers = tl
    sys.modules.update({"timm": timm, "timm.models": tm, "timm.models.layers": tl})

    from train_utils.model_selection import init_loss_func          # the reference's own dispatch, unmodified
    import models.loss, models.FOCALModules, models.DeepSense
    import focal_b200.loss
    assert models.loss.__file__.startswith(root + "/focal_b200/dropin"), models.loss.__file__
    assert models.FOCALModules.__file__.startswith(ref) and models.DeepSense.__file__.startswith(ref)
    args = types.SimpleNamespace(
        device="cpu", model="DeepSense", tag=None, train_mode="contrastive", stage="pretrain", learn_framework="FOCAL",
        dataset_config={"modality_names": ["seismic", "audio"], "seq_len": 4,
                        "FOCAL": {"temperature": {"DeepSense": 0.5, "SW_Transformer": 0.07}, "inter_rank_margin": 1,
                                  "shared_contrastive_loss_weight": 1, "private_contrastive_loss_weight": 1,
                                  "orthogonal_loss_weight": 3, "rank_loss_weight": 5}})
    loss_func = init_loss_func(args)
    assert type(loss_func) is focal_b200.loss.FOCALLoss, type(loss_func)
    assert len(loss_func.state_dict()) == 0
    print("DROPIN_OK")
''')


@pytest.mark.skipif(not os.path.isdir(REF), reason="/root/reference only exists in the build container")
def test_reference_dispatch_resolves_to_our_loss():
    res = subprocess.run([sys.executable, "-c", SCRIPT, ROOT, REF], capture_output=True, text=True, timeout=300)
    assert res.returncode == 0 and "DROPIN_OK" in res.stdout, res.stdout + res.stderr
