"""One rank's share of the row-sharded step on ONE GPU (world = 1 peer table, owned sequences = rank r of R): what the
kernels of a rank cost without any cross-GPU wait.  Run under `ncu --metrics gpu__time_duration.sum` for the
per-kernel split, or plain for the CUDA-event total.

    python tools/shard_probe.py 3 8            # rank 3 of 8 of the headline batch
"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from focal_b200 import _cabi
from focal_b200.engine import CudaBackend, FocalHyper, shard_sequences


def main():
    r, R = int(sys.argv[1]), int(sys.argv[2])
    B, D, S, mods = 8192, 256, 4, ("seismic", "audio")
    steps = int(sys.argv[3]) if len(sys.argv) > 3 else 20
    hp = FocalHyper(mods, S, 0.5, 1.0, 1.0, 1.0, 3.0, 5.0)
    be = CudaBackend()
    lib = be.lib
    cfg = be._cfg(hp, B, D, True, shard_sequences(B // S, R, r))
    cfg.local_rows = 1
    info = _cabi.FocalWsInfo()
    _cabi.check(lib.focal_b200_workspace_info(C.byref(cfg), C.byref(info)), "workspace_info")
    ws, handle = C.c_void_p(), C.create_string_buffer(64)
    _cabi.check(lib.focal_b200_peer_alloc(info.total_bytes, C.byref(ws), handle), "peer_alloc")
    peers = _cabi.FocalPeers(rank=0, world=1)
    peers.ws[0] = ws.value
    Bl = B // R
    torch.manual_seed(0)
    local = [torch.randn(Bl, D, device="cuda") for _ in range(4)]
    grads = [torch.empty_like(t) for t in local]
    loss5 = torch.empty(5, device="cuda")
    fptr, gptr = _cabi.ptr_array([t.data_ptr() for t in local]), _cabi.ptr_array([g.data_ptr() for g in grads])
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)

    def step():
        _cabi.check(lib.focal_b200_loss_sharded(C.byref(cfg), fptr, C.byref(peers), C.c_size_t(info.total_bytes),
                                                C.c_void_p(loss5.data_ptr()), gptr, st), "loss_sharded")
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        step()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    g.replay()
    torch.cuda.synchronize()
    e0.record()
    for _ in range(steps):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    print(f"rank {r}/{R}: {e0.elapsed_time(e1) / steps * 1e3:.1f} us per step (graph replay, no peers), "
          f"pieces nce={info.n_pieces_nce} tmp={info.n_pieces_tmp}")
    torch.cuda.synchronize()
    lib.focal_b200_peer_free(ws)


if __name__ == "__main__":
    main()
