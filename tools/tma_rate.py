"""Micro-benchmark (GPU): per-SM throughput of linear TMA bulk copies (cp.async.bulk) from L2 into shared memory."""
import ctypes as C, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from focal_b200 import _cabi
lib = _cabi.load_bringup()
span = 32 << 20                                   # 32 MiB source: L2 resident
src = torch.zeros(span + (1 << 20), dtype=torch.uint8, device="cuda")
for grid in (1, 148):
    for copy_bytes, per_stage, stages in ((16384, 2, 5), (16384, 2, 2), (8192, 4, 4), (12288, 4, 3), (32768, 1, 5),
                                          (4096, 8, 5), (16384, 1, 8), (16384, 4, 3)):
        out = torch.zeros(grid * 2, dtype=torch.int64, device="cuda")
        iters = 400
        rc = lib.focal_b200_debug_tma_rate(C.c_void_p(src.data_ptr()), span, copy_bytes, per_stage, stages, iters, grid,
                                           C.c_void_p(out.data_ptr()), None)
        assert rc == 0
        torch.cuda.synchronize()
        cyc = out.view(grid, 2)[:, 1].float().mean().item()
        stage_bytes = copy_bytes * per_stage
        print(f"grid={grid:3d} copy={copy_bytes:6d} B x{per_stage} per stage, {stages} stages: "
              f"{cyc / iters:7.1f} clk/stage  {stage_bytes * iters / cyc:6.1f} B/clk/SM")
