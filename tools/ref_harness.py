"""Run the reference's own training code on the B200 with `models.loss` shadowed by focal_b200.

TEST / BENCH INFRASTRUCTURE.  Uses the unmodified reference tree that `python -m oracle.build_ref` placed under
oracle/_ref/src (or the live /root/reference/src in the build container) and:

* `make_run_dir`  -- a scratch run directory: a copy of the reference `src/` (it reads `./data/MOD.yaml` and writes
  `../weights` relative to the cwd, params/params_util.py:119, params/output_paths.py:99), a tiny synthetic MOD-format
  dataset (`<seq>_<idx>.pt` files holding {"data": {"shake": {"audio": [1, 10, 1600], "seismic": [1, 10, 20]}},
  "label": {...}}, input_utils/multi_modal_dataset.py:39-56,80-93), index `.txt` files, and the YAML pointed at them
  with the epoch count cut down;
* `run_train_literal` -- `runpy` of the reference's `train.py` (unchanged) in that directory with
  `focal_b200/dropin` ahead of it on `sys.path`, i.e. `python train.py -model=DeepSense -learn_framework=FOCAL`;
* `build_pretrain_objects` -- args / backbone / augmenter / loss function built by the reference's own
  `init_backbone_model`, `init_pretrain_framework`, `Augmenter`, `init_loss_func`, for step-level tests and timing of
  `calc_pretrain_loss` (train_utils/loss_calc_utils.py:1-22).
"""
from __future__ import annotations

import os
import runpy
import shutil
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

REF_PACKAGES = ("models", "general_utils", "train_utils", "input_utils", "data_augmenter", "params")


def make_run_dir(tmp: str, n_seq: int = 16, samples_per_seq: int = 8, epochs: int = 1, seed: int = 0,
                 model: str = "DeepSense") -> str:
    """Returns the directory to run `train.py` from (<tmp>/src)."""
    import torch
    import yaml

    from oracle.build_ref import reference_sys_path
    src = os.path.join(tmp, "src")
    shutil.copytree(reference_sys_path(), src)
    data_dir = os.path.join(tmp, "synthetic_mod")
    os.makedirs(data_dir)
    g = torch.Generator().manual_seed(seed)
    files = []
    for s in range(n_seq):
        base = torch.randn(1, 10, 1600, generator=g), torch.randn(1, 10, 20, generator=g)
        for i in range(samples_per_seq):
            sample = {
                "data": {"shake": {"audio": base[0] + 0.3 * torch.randn(1, 10, 1600, generator=g),
                                   "seismic": base[1] + 0.3 * torch.randn(1, 10, 20, generator=g)}},
                "label": {"vehicle_type": torch.tensor(s % 7), "distance": torch.tensor(s % 3),
                          "speed": torch.tensor(s % 4)},
            }
            path = os.path.join(data_dir, f"run{s:03d}_{i}.pt")       # sequence id = everything before the last "_"
            torch.save(sample, path)
            files.append(path)
    idx_dir = os.path.join(tmp, "index")
    os.makedirs(idx_dir)
    index = {}
    for name in ("pretrain", "train", "val", "test"):
        index[name] = os.path.join(idx_dir, f"{name}_index.txt")
        with open(index[name], "w") as fh:
            fh.write("\n".join(files) + "\n")
    ypath = os.path.join(src, "data", "MOD.yaml")
    with open(ypath) as fh:
        cfg = yaml.safe_load(fh)
    cfg["pretrain_index_file"] = index["pretrain"]
    for task in ("vehicle_classification", "distance_classification", "speed_classification"):
        cfg[task]["train_index_file"] = index["train"]
        cfg[task]["val_index_file"] = index["val"]
        cfg[task]["test_index_file"] = index["test"]
    cfg["FOCAL"]["pretrain_lr_scheduler"]["train_epochs"] = epochs
    with open(ypath, "w") as fh:
        yaml.safe_dump(cfg, fh)
    return src


class _ReferenceEnv:
    """sys.path / sys.modules / cwd set up like `cd src; PYTHONPATH=<repo>:<repo>/focal_b200/dropin python ...`."""

    def __init__(self, src: str, shadow_loss: bool):
        self.src, self.shadow = src, shadow_loss

    def __enter__(self):
        from oracle.build_ref import install_import_stubs
        install_import_stubs()
        self.saved_path, self.saved_cwd, self.saved_argv = list(sys.path), os.getcwd(), list(sys.argv)
        self.saved_mods = {k: v for k, v in sys.modules.items() if k.split(".")[0] in REF_PACKAGES}
        for k in self.saved_mods:
            del sys.modules[k]
        head = [ROOT] + ([os.path.join(ROOT, "focal_b200", "dropin")] if self.shadow else []) + [self.src]
        sys.path[:0] = head
        os.chdir(self.src)
        return self

    def __exit__(self, *exc):
        os.chdir(self.saved_cwd)
        sys.argv[:] = self.saved_argv
        sys.path[:] = self.saved_path
        for k in [k for k in sys.modules if k.split(".")[0] in REF_PACKAGES]:
            del sys.modules[k]
        sys.modules.update(self.saved_mods)
        return False


def run_train_literal(src: str, extra_argv=(), shadow_loss: bool = True, workers: int = 0) -> dict:
    """`python train.py -model=DeepSense -dataset=MOD -learn_framework=FOCAL ...` from <src>, in this process."""
    argv = ["train.py", "-model=DeepSense", "-dataset=MOD", "-learn_framework=FOCAL", "-gpu=0", *extra_argv]
    with _ReferenceEnv(src, shadow_loss):
        sys.argv[:] = argv
        if workers is not None:
            # the reference hard-codes 10 DataLoader workers (params_util.py:132); the tiny synthetic set does not need them
            import params.params_util as pu
            orig = pu.set_auto_params

            def patched(args):
                args = orig(args)
                args.workers = workers
                return args
            pu.set_auto_params = patched
            import params.train_params as tp
            tp.set_auto_params = patched
        ns = runpy.run_path(os.path.join(src, "train.py"), run_name="__main__")
        import models.loss as ml
        return {"loss_module": ml.__file__, "loss_class": ml.FOCALLoss, "globals": ns}


def build_pretrain_objects(src: str, batch_size: int, model: str = "DeepSense", shadow_loss: bool = True,
                           precision: str = "auto"):
    """Everything `pretrain()` builds before its loop, from the reference's own factories.  Returns a namespace with
    args, default_model, augmenter, loss_func, calc_pretrain_loss, optimizer, env (keep `env` entered while using them)."""
    import torch
    env = _ReferenceEnv(src, shadow_loss)
    env.__enter__()
    from input_utils.yaml_utils import load_yaml                          # type: ignore
    args = types.SimpleNamespace(
        tag=None, dataset="MOD", task="vehicle_classification", model=model, learn_framework="FOCAL", stage="pretrain",
        label_ratio=1.0, model_weight=None, batch_size=batch_size, gpu="0", option="train",
        device=torch.device("cuda", 0), half=False, train_mode="contrastive", sequence_sampler=True, workers=0,
        focal_precision=precision)
    args.dataset_config = load_yaml(os.path.join(src, "data", "MOD.yaml"))
    from data_augmenter.Augmenter import Augmenter                        # type: ignore
    from train_utils.loss_calc_utils import calc_pretrain_loss            # type: ignore
    from train_utils.model_selection import (init_backbone_model, init_loss_func,  # type: ignore
                                             init_pretrain_framework)
    from train_utils.optimizer import define_optimizer                    # type: ignore
    augmenter = Augmenter(args)
    augmenter.to(args.device)
    args.augmenter = augmenter
    backbone = init_backbone_model(args)
    args.classifier = backbone
    loss_func = init_loss_func(args)
    default_model = init_pretrain_framework(args, backbone)
    optimizer = define_optimizer(args, default_model.parameters())
    return types.SimpleNamespace(args=args, default_model=default_model, augmenter=augmenter, loss_func=loss_func,
                                 calc_pretrain_loss=calc_pretrain_loss, optimizer=optimizer, env=env)


def synthetic_time_inputs(batch: int, seed: int = 0, device="cuda"):
    """{loc: {mod: [B, c, intervals, spectrum]}} time-domain batch in the MOD shapes (MOD.yaml:33-52)."""
    import torch
    g = torch.Generator().manual_seed(seed)
    return {"shake": {"audio": torch.randn(batch, 1, 10, 1600, generator=g).to(device),
                      "seismic": torch.randn(batch, 1, 10, 20, generator=g).to(device)}}
