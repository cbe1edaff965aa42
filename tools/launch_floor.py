"""Device-side cost of a launch as a function of what the kernel asks for (tools/launch_floor.py on the GPU box):
20 launches captured in a CUDA graph, replayed 20 times, CUDA events around the replays."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from focal_b200._cabi import load_bringup

lib = load_bringup()
lib.focal_b200_debug_launch_floor.argtypes = [C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p]
lib.focal_b200_debug_launch_floor.restype = C.c_int
out = torch.zeros(4096, dtype=torch.int32, device="cuda")
names = {1: "10.7 KB params", 2: "200 KB smem", 4: "TMEM alloc", 8: "row launch between"}
for flags in (0, 1, 2, 4, 3, 7, 8, 9, 10, 15):
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        assert lib.focal_b200_debug_launch_floor(flags, 3, out.data_ptr(), s.cuda_stream) == 0
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            assert lib.focal_b200_debug_launch_floor(flags, 20, out.data_ptr(), s.cuda_stream) == 0
        for _ in range(3):
            g.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            g.replay()
        e1.record()
        torch.cuda.synchronize()
    per = e0.elapsed_time(e1) * 1e3 / 400
    what = " + ".join(v for k, v in names.items() if flags & k) or "empty 148 x 576 launch"
    print(f"flags {flags:2d}: {per:6.2f} us per iteration   ({what})")
