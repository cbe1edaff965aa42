"""Recipe that puts the UNMODIFIED reference where the GPU box can run it: ``oracle/_ref/``.

TEST / BENCH INFRASTRUCTURE ONLY -- see ``oracle/__init__.py``.

The reference (tomoyoshki/focal) is pure Python; "building" it means taking its source tree as it lies under
``/root/reference/src`` and placing a byte-identical copy under ``oracle/_ref/src`` (git-ignored build output, NOT
gpurun-ignored, so it travels to the GPU box like a built ``.so``; nothing of it is ever committed).  Beside the copy
go import-time stubs for the three third-party packages the reference imports but this image lacks (``matplotlib``,
``timm``, ``tsai``: SURVEY.md section 8c) and a manifest with the SHA-256 of every copied file.

Consumers (tests/, bench.py's reference arm and cpu_baseline leg, smoke()):
    from oracle.build_ref import ref_available, import_reference_loss, reference_sys_path

    python -m oracle.build_ref            # (re)build; a no-op without /root/reference
"""
from __future__ import annotations

import contextlib
import hashlib
import json
import os
import shutil
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = "/root/reference/src"
OUT = os.path.join(HERE, "_ref")
OUT_SRC = os.path.join(OUT, "src")
MANIFEST = os.path.join(OUT, "MANIFEST.json")


def _sha(path: str) -> str:
    h = hashlib.sha256()
    with open(path, "rb") as fh:
        h.update(fh.read())
    return h.hexdigest()


def build(force: bool = False) -> bool:
    """Copy /root/reference/src -> oracle/_ref/src (Python sources + data/*.yaml).  Returns True when oracle/_ref is
    usable afterwards (freshly built or already there)."""
    if not os.path.isdir(REF_SRC):
        return ref_available()
    if ref_available() and not force:
        with open(MANIFEST) as fh:
            man = json.load(fh)
        if all(os.path.exists(os.path.join(REF_SRC, rel)) and _sha(os.path.join(REF_SRC, rel)) == dig
               for rel, dig in man["files"].items()):
            return True
    if os.path.isdir(OUT):
        shutil.rmtree(OUT)
    files = {}
    for root, _dirs, names in os.walk(REF_SRC):
        for n in names:
            if not n.endswith((".py", ".yaml", ".yml", ".txt")):
                continue
            src = os.path.join(root, n)
            rel = os.path.relpath(src, REF_SRC)
            dst = os.path.join(OUT_SRC, rel)
            os.makedirs(os.path.dirname(dst), exist_ok=True)
            shutil.copyfile(src, dst)
            files[rel] = _sha(dst)
    with open(MANIFEST, "w") as fh:
        json.dump({"source": REF_SRC, "files": files}, fh, indent=1, sort_keys=True)
    return True


def ref_available() -> bool:
    return os.path.exists(MANIFEST) and os.path.exists(os.path.join(OUT_SRC, "models", "loss.py"))


def install_import_stubs() -> None:
    """matplotlib / timm / tsai are imported by reference modules beyond models.loss but are absent from this image
    (SURVEY.md section 8c).  The stubs provide exactly the names those imports ask for; none is on the loss path."""
    import torch.nn as nn

    def mod(name):
        if name in sys.modules:
            return sys.modules[name]
        m = types.ModuleType(name)
        sys.modules[name] = m
        parent, _, child = name.rpartition(".")
        if parent:
            setattr(mod(parent), child, m)
        return m

    try:
        import matplotlib.pyplot  # noqa: F401
    except Exception:
        mod("matplotlib")
        mod("matplotlib.pyplot").axis = lambda *a, **k: None
    try:
        import timm  # noqa: F401
    except Exception:
        layers = mod("timm.models.layers")

        class DropPath(nn.Module):
            def __init__(self, drop_prob=0.0):
                super().__init__()
                self.drop_prob = drop_prob

            def forward(self, x):
                return x

        def trunc_normal_(tensor, mean=0.0, std=1.0, a=-2.0, b=2.0):
            return nn.init.trunc_normal_(tensor, mean=mean, std=std, a=a, b=b)

        def to_2tuple(x):
            return tuple(x) if isinstance(x, (tuple, list)) else (x, x)

        layers.DropPath, layers.trunc_normal_, layers.to_2tuple = DropPath, trunc_normal_, to_2tuple
        mod("timm.models").layers = layers
        sched = mod("timm.scheduler")

        class _Sched:
            def __init__(self, optimizer, *a, **k):
                self.optimizer = optimizer

            def step(self, *a, **k):
                pass

            def step_update(self, *a, **k):
                pass

        for nm in ("cosine_lr", "step_lr", "scheduler", "multistep_lr"):
            mod(f"timm.scheduler.{nm}")
        sys.modules["timm.scheduler.cosine_lr"].CosineLRScheduler = _Sched
        sys.modules["timm.scheduler.step_lr"].StepLRScheduler = _Sched
        sys.modules["timm.scheduler.multistep_lr"].MultiStepLRScheduler = _Sched
        sys.modules["timm.scheduler.scheduler"].Scheduler = _Sched
        sched.CosineLRScheduler = sched.StepLRScheduler = sched.MultiStepLRScheduler = sched.Scheduler = _Sched
    try:
        import tsai  # noqa: F401
    except Exception:
        tr = mod("tsai.data.transforms")
        core = mod("tsai.data.core")

        class _Identity:
            def __init__(self, *a, **k):
                pass

            def __call__(self, x, *a, **k):
                return x

        for nm in ("TSTimeWarp", "TSMagWarp", "TSMagScale", "TSTimeNoise", "TSRandomShift", "TSHorizontalFlip"):
            setattr(tr, nm, _Identity)
        core.TSTensor = lambda x, *a, **k: x
        mod("tsai.data").transforms = tr
        mod("tsai").data = sys.modules["tsai.data"]


def reference_sys_path() -> str:
    """Directory to put on sys.path to import the reference (oracle/_ref/src, else the live /root/reference/src)."""
    if ref_available():
        return OUT_SRC
    if os.path.isdir(REF_SRC):
        return REF_SRC
    raise ImportError("the reference is not available: run `python -m oracle.build_ref` where /root/reference exists")


@contextlib.contextmanager
def reference_on_path():
    """Temporarily import `models`, `general_utils`, ... from the reference (and not from focal_b200/dropin)."""
    path = reference_sys_path()
    saved_path = list(sys.path)
    saved_mods = {k: v for k, v in sys.modules.items()
                  if k.split(".")[0] in ("models", "general_utils", "train_utils", "input_utils", "data_augmenter", "params")}
    for k in saved_mods:
        del sys.modules[k]
    sys.path.insert(0, path)
    try:
        yield path
    finally:
        sys.path[:] = saved_path
        for k in [k for k in sys.modules
                  if k.split(".")[0] in ("models", "general_utils", "train_utils", "input_utils", "data_augmenter", "params")]:
            del sys.modules[k]
        sys.modules.update(saved_mods)


def import_reference_loss():
    """The reference's own ``FOCALLoss`` class (src/models/loss.py:8), unmodified."""
    with reference_on_path():
        from models.loss import FOCALLoss  # type: ignore
        return FOCALLoss


if __name__ == "__main__":
    ok = build(force="--force" in sys.argv)
    print("oracle/_ref:", "ready" if ok else "not built (no /root/reference here)")
