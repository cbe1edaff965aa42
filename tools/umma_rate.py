"""Micro-benchmark (GPU): cycles per tcgen05.mma M=128 for several N, SS mode, bf16."""
import ctypes as C, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from focal_b200 import _cabi
lib = C.CDLL(_cabi.LIB_PATH)
lib.focal_b200_debug_umma_rate.argtypes = [C.c_uint32] * 5 + [C.c_void_p, C.c_void_p]
for grid in (148,):
  for a_tmem in (0, 1):
    for b_mn in (0, 1):
        for N in ((32, 64, 96, 128, 192, 256) if not a_tmem else (32, 64, 96, 128)):
          if True:
            ks = 8
            out = torch.zeros(grid, dtype=torch.int64, device="cuda")
            iters = 2000
            rc = lib.focal_b200_debug_umma_rate(N, b_mn | (a_tmem << 1), iters, ks, grid, C.c_void_p(out.data_ptr()), None)
            torch.cuda.synchronize()
            cyc = out.float().mean().item() / (iters * ks)
            print(f"a_tmem={a_tmem} grid={grid:3d} b_mn={b_mn} N={N:3d} K=16: {cyc:7.1f} clk/mma  (MACs/clk/SM = {128*N*16/cyc:7.0f})")
