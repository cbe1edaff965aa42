"""GPU diagnostic: host enqueue time of each of the first steps through FocalEngine (eager, capture, replays)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from focal_b200.engine import FocalEngine, FocalHyper

B, D = int(os.environ.get("FB_B", 8192)), int(os.environ.get("FB_D", 256))
mods = ("seismic", "audio")
hp = FocalHyper(mods, 4, 0.5, 1.0, 1.0, 1.0, 3.0, 5.0, False, 7, os.environ.get("FB_PREC", "bf16"))
eng = FocalEngine(hp)
sets = [[torch.randn(B, D, device="cuda") for _ in range(4)] for _ in range(8)]
torch.cuda.synchronize()
host, evs = [], []
for k in range(40):
    x = sets[k % 8]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    out = eng.loss_and_grads({m: x[i] for i, m in enumerate(mods)}, {m: x[2 + i] for i, m in enumerate(mods)}, True)
    e1.record()
    host.append((time.perf_counter() - t0) * 1e3)
    evs.append((e0, e1))
torch.cuda.synchronize()
print("step: host ms / device ms")
print(" ".join(f"{k}:{h:.3f}/{a.elapsed_time(b):.3f}" for k, (h, (a, b)) in enumerate(zip(host, evs))))
