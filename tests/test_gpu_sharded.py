"""GPU: the row-sharded peer-memory path (focal_b200_loss_sharded) with all ranks emulated on ONE device.

Kept in its own module, collected after the other GPU tests: the ranks wait for each other inside kernels, so a protocol
bug shows up as a device-side trap (bounded spin), which poisons the CUDA context for whatever runs afterwards.
"""
import pytest
import torch

from oracle import focal_oracle as fo

pytestmark = pytest.mark.gpu


def _require_cuda():
    assert torch.cuda.is_available(), "GPU tests selected (-m gpu) but no CUDA device is visible"


@pytest.mark.parametrize("world,B,D,mods,T,need_grad,S", [
    (1, 512, 256, ("seismic", "audio"), 0.5, True, 4),
    (2, 2048, 256, ("seismic", "audio"), 0.5, True, 4),
    (4, 1536, 128, ("acc", "gyr", "mag"), 0.07, True, 4),
    (2, 1024, 64, ("seismic", "audio"), 0.5, False, 2),
    (2, 1024, 128, ("seismic", "audio"), 0.5, True, 1),       # "global" InfoNCE (cfg 4 shape): temporal term is NaN
    (2, 1024, 512, ("seismic", "audio"), 0.5, True, 4),       # wide temporal mode (8 K blocks), vectorised rows VW = 8
])
def test_sharded_peer_path_on_one_gpu(world, B, D, mods, T, need_grad, S):
    """focal_b200_loss_sharded (prologue of the owned rows storing into every rank's workspace, device-side
    barriers, in-kernel loss all-reduce) with all `world` ranks emulated on ONE GPU: one workspace and one stream per
    rank, the barrier kernels of the ranks spin concurrently.  Every rank must report the single-GPU loss and the
    single-GPU gradients of its rows; repeated steps keep the barrier epochs in step."""
    _require_cuda()
    import ctypes as C
    from focal_b200 import _cabi
    from focal_b200.engine import CudaBackend, FocalHyper, shard_sequences
    f1, f2 = fo.make_structured(11, list(mods), B, D, S)
    hp = FocalHyper(tuple(mods), S, T, 1.0, 1.0, 1.0, 3.0, 5.0)
    be = CudaBackend()
    feats = [f1[m].cuda() for m in mods] + [f2[m].cuda() for m in mods]
    b = B // S
    want5, wantg = be.run(hp, feats, (0, b), need_grad, None)
    torch.cuda.synchronize()
    lib = be.lib
    cfgs = []
    for r in range(world):
        c = be._cfg(hp, B, D, need_grad, shard_sequences(b, world, r))
        c.local_rows = 1
        cfgs.append(c)
    info = _cabi.FocalWsInfo()
    _cabi.check(lib.focal_b200_workspace_info(C.byref(cfgs[0]), C.byref(info)), "workspace_info")
    wss = []
    try:
        for r in range(world):
            ptr, handle = C.c_void_p(), C.create_string_buffer(64)
            _cabi.check(lib.focal_b200_peer_alloc(info.total_bytes, C.byref(ptr), handle), "peer_alloc")
            wss.append(ptr)
        Bl = B // world
        local = [[t[r * Bl:(r + 1) * Bl].contiguous() for t in feats] for r in range(world)]
        streams = [torch.cuda.Stream() for _ in range(world)]
        # Load every kernel this shard shape uses BEFORE ranks start waiting for each other: with CUDA's lazy module
        # loading the first launch of a kernel can block the host until running kernels finish, and here (one process
        # driving all ranks) a rank's spinning wait only finishes once the host has launched the other ranks.
        ptr, handle = C.c_void_p(), C.create_string_buffer(64)
        _cabi.check(lib.focal_b200_peer_alloc(info.total_bytes, C.byref(ptr), handle), "peer_alloc")
        solo = _cabi.FocalPeers(rank=0, world=1)
        solo.ws[0] = ptr.value
        l5 = torch.empty(5, device="cuda")
        gw = [torch.empty_like(t) for t in local[0]]
        c0 = be._cfg(hp, B, D, need_grad, shard_sequences(b, world, 0))
        c0.local_rows = 1
        _cabi.check(lib.focal_b200_loss_sharded(C.byref(c0), _cabi.ptr_array([t.data_ptr() for t in local[0]]),
                                                C.byref(solo), C.c_size_t(info.total_bytes), C.c_void_p(l5.data_ptr()),
                                                _cabi.ptr_array([g.data_ptr() for g in gw]) if need_grad else None,
                                                C.c_void_p(torch.cuda.current_stream().cuda_stream)), "warm-up")
        torch.cuda.synchronize()
        lib.focal_b200_peer_free(ptr)
        for step in range(3):
            outs = []
            for r in range(world):
                loss5 = torch.full((5,), float("nan"), device="cuda")
                grads = [torch.full_like(t, float("nan")) for t in local[r]] if need_grad else None
                outs.append((loss5, grads))
            torch.cuda.synchronize()
            for r in range(world):         # no host sync between the ranks: their barrier kernels wait for each other
                peers = _cabi.FocalPeers(rank=r, world=world)
                for q in range(world):
                    peers.ws[q] = wss[q].value
                loss5, grads = outs[r]
                gptr = _cabi.ptr_array([g.data_ptr() for g in grads]) if need_grad else None
                rc = lib.focal_b200_loss_sharded(C.byref(cfgs[r]), _cabi.ptr_array([t.data_ptr() for t in local[r]]),
                                                 C.byref(peers), C.c_size_t(info.total_bytes),
                                                 C.c_void_p(loss5.data_ptr()), gptr, C.c_void_p(streams[r].cuda_stream))
                _cabi.check(rc, "focal_b200_loss_sharded")
            torch.cuda.synchronize()
            for r, (loss5, grads) in enumerate(outs):
                assert torch.allclose(loss5, want5, rtol=2e-6, atol=1e-6, equal_nan=True), \
                    (step, r, loss5.tolist(), want5.tolist())
                if need_grad:
                    for g, wg in zip(grads, wantg):
                        ref = wg[r * Bl:(r + 1) * Bl]
                        assert float((g - ref).norm() / ref.norm()) < 1e-5, (step, r)
    finally:
        torch.cuda.synchronize()
        for p in wss:
            lib.focal_b200_peer_free(p)


def test_sharded_rejects_bad_arguments():
    _require_cuda()
    import ctypes as C
    from focal_b200 import _cabi
    from focal_b200.engine import CudaBackend, FocalHyper
    be = CudaBackend()
    hp = FocalHyper(("a", "b"), 4, 0.5, 1.0, 1.0, 1.0, 3.0, 5.0)
    cfg = be._cfg(hp, 512, 256, True, (0, 128))
    peers = _cabi.FocalPeers(rank=0, world=1)
    loss5 = torch.empty(5, device="cuda")
    x = [torch.randn(512, 256, device="cuda") for _ in range(4)]
    fptr = _cabi.ptr_array([t.data_ptr() for t in x])
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    # local_rows not set / no workspace / staged entry points refuse local_rows
    assert be.lib.focal_b200_loss_sharded(C.byref(cfg), fptr, C.byref(peers), C.c_size_t(1 << 30),
                                          C.c_void_p(loss5.data_ptr()), fptr, st) == _cabi.FOCAL_EINVAL
    cfg.local_rows = 1
    assert be.lib.focal_b200_loss_sharded(C.byref(cfg), fptr, C.byref(peers), C.c_size_t(1 << 30),
                                          C.c_void_p(loss5.data_ptr()), fptr, st) == _cabi.FOCAL_EINVAL
    ws, info = be.workspace(be._cfg(hp, 512, 256, True, (0, 128)), x[0].device)
    assert be.lib.focal_b200_prologue(C.byref(cfg), fptr, C.c_void_p(ws.data_ptr()), C.c_size_t(ws.numel()),
                                      st) == _cabi.FOCAL_EINVAL
    # shapes off the vectorised row-kernel path are refused (callers fall back to the collective path)
    hp8 = FocalHyper(("a", "b"), 8, 0.5, 1.0, 1.0, 1.0, 3.0, 5.0)
    assert not CudaBackend.peer_eligible(hp8, 256, 2) and CudaBackend.peer_eligible(hp, 256, 8)
