"""Micro-benchmark (GPU): tensor-pipe clocks per column tile of the fused temporal kernel's MMA schedule, no epilogue."""
import ctypes as C, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from focal_b200 import _cabi
lib = _cabi.load_bringup()
lib.focal_b200_debug_umma_tile_rate.argtypes = [C.c_uint32] * 4 + [C.c_void_p, C.c_void_p]
lib.focal_b200_debug_umma_tile_rate.restype = C.c_int
grid, iters = 148, 2000
names = {0: "alternating #1/#2 (kernel order)", 1: "only UMMA #1 (SS, N=BN)", 2: "only UMMA #2 (TS, N=256)", 3: "two tiles batched"}
for BN in (64, 80, 96, 128):
    for mode in (0, 1, 2, 3):
        out = torch.zeros(grid, dtype=torch.int64, device="cuda")
        rc = lib.focal_b200_debug_umma_tile_rate(BN, mode, iters, grid, C.c_void_p(out.data_ptr()), None)
        assert rc == 0, rc
        torch.cuda.synchronize()
        per_tile = out.float().mean().item() / iters
        print(f"BN={BN:3d} {names[mode]:34s}: {per_tile:7.1f} clk per tile = {per_tile / BN:5.2f} clk per column "
              f"(nominal tensor time {16 * BN // 2 + (BN // 16) * 128} clk)")
