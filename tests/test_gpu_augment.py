"""GPU: the frequency-domain input stage (focal_b200.augment) against the reference's own
``Augmenter.fft_preprocess`` and ``PhaseShiftAugmenter`` (unmodified, from oracle/_ref) on the same inputs and the same
``random`` stream."""
import math
import os
import random
import types

import pytest
import torch

pytestmark = pytest.mark.gpu


def _args(dev, p=1.0):
    return types.SimpleNamespace(device=dev, dataset_config={
        "modality_names": ["seismic", "audio"], "location_names": ["shake"], "phase_shift": {"prob": p}})


def _inputs(b, dev, seed=0):
    g = torch.Generator().manual_seed(seed)
    return {"shake": {"audio": torch.randn(b, 1, 10, 1600, generator=g).to(dev),
                      "seismic": torch.randn(b, 1, 10, 20, generator=g).to(dev)}}


def _reference_classes():
    from oracle.build_ref import install_import_stubs, reference_on_path, ref_available
    if not ref_available() and not os.path.isdir("/root/reference/src"):
        pytest.fail("oracle/_ref is missing: run `python -m oracle.build_ref` where /root/reference exists")
    install_import_stubs()
    with reference_on_path():
        from data_augmenter.Augmenter import Augmenter                    # type: ignore
        from data_augmenter.PhaseShiftAugmenter import PhaseShiftAugmenter  # type: ignore
        return Augmenter, PhaseShiftAugmenter


def test_fft_preprocess_matches_the_reference_bitwise():
    assert torch.cuda.is_available()
    from focal_b200 import augment
    Augmenter, _ = _reference_classes()
    dev = torch.device("cuda", 0)
    x = _inputs(64, dev)
    want = Augmenter.fft_preprocess(None, x)                              # the method does not touch self
    got = augment.fft_preprocess(x)
    for mod in ("audio", "seismic"):
        assert got["shake"][mod].shape == want["shake"][mod].shape
        assert torch.equal(got["shake"][mod], want["shake"][mod].contiguous()), mod   # same cuFFT, pure layout change


@pytest.mark.parametrize("p", [1.0, 0.5])
def test_phase_shift_matches_the_reference_with_the_same_random_stream(p):
    assert torch.cuda.is_available()
    from focal_b200 import augment
    Augmenter, RefPhase = _reference_classes()
    dev = torch.device("cuda", 0)
    freq = augment.fft_preprocess(_inputs(32, dev, seed=1))
    ref_aug, our_aug = RefPhase(_args(dev, p)), augment.PhaseShiftAugmenter(_args(dev, p))
    for trial in range(4):
        random.seed(100 + trial)
        want, want_lab, _ = ref_aug(freq, None)
        state_after_ref = random.getstate()
        random.seed(100 + trial)
        got, got_lab, _ = our_aug(freq, None)
        assert random.getstate() == state_after_ref                       # same number of draws, same decisions
        assert torch.equal(got_lab, want_lab)
        for mod in ("audio", "seismic"):
            a, b = got["shake"][mod], want["shake"][mod]
            scale = float(b.abs().max())
            # the reference goes through abs / angle / cos / sin in fp32; a direct rotation differs at rounding level
            assert float((a - b).abs().max()) <= 2e-5 * scale, (trial, mod, float((a - b).abs().max()), scale)


def test_fused_rotation_equals_layout_change_then_phase_shift_and_rejects_cpu():
    assert torch.cuda.is_available()
    from focal_b200 import augment
    dev = torch.device("cuda", 0)
    x = _inputs(8, dev, seed=2)["shake"]["audio"]
    spec = torch.fft.fft(x, dim=-1)
    ang = 0.7
    fused = augment.spectrum_to_channels(spec, ang)
    two_step = augment.phase_shift(augment.spectrum_to_channels(spec), ang)
    assert torch.equal(fused, two_step)
    rot = spec * complex(math.cos(ang), math.sin(ang))
    want = torch.view_as_real(rot).permute(0, 1, 4, 2, 3).reshape(fused.shape)
    assert float((fused - want).abs().max()) <= 1e-5 * float(want.abs().max())
    with pytest.raises(RuntimeError):
        augment.phase_shift(torch.randn(2, 2, 10, 20), 0.1)
