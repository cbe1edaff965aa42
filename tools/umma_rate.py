"""Micro-benchmark (GPU): cycles per tcgen05.mma (M=128, bf16) vs N / operand placement / per-tile barrier traffic."""
import ctypes as C, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from focal_b200 import _cabi
lib = _cabi.load_bringup()
grid = 148
for sync_mode in (0, 1, 3):
    for a_tmem, b_mn, N in ((0, 0, 64), (0, 0, 128), (0, 0, 256), (1, 1, 128)):
        out = torch.zeros(grid, dtype=torch.int64, device="cuda")
        iters = 2000
        rc = lib.focal_b200_debug_umma_rate(N, b_mn | (a_tmem << 1), iters, sync_mode, grid, C.c_void_p(out.data_ptr()), None)
        assert rc == 0
        torch.cuda.synchronize()
        per_iter = out.float().mean().item() / iters
        print(f"sync_mode={sync_mode} a_tmem={a_tmem} b_mn={b_mn} N={N:3d}: {per_iter:7.1f} clk per 8 MMAs "
              f"({per_iter / 8:6.1f} clk/mma; tensor-pipe time {8 * N / 2} clk)")
