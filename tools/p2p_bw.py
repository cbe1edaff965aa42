"""Peer-to-peer bandwidth between two GPUs of the box: copy engine (tensor.copy_ across devices) and SM stores (a
torch elementwise kernel writing into peer memory), for the message sizes of the row-sharded operand exchange."""
import torch
assert torch.cuda.device_count() >= 2
for mb in (4, 16, 64):
    n = mb * 2 ** 20 // 4
    src = torch.randn(n, device="cuda:0")
    dst = torch.empty(n, device="cuda:1")
    torch.cuda.synchronize(0); torch.cuda.synchronize(1)
    with torch.cuda.device(0):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for _ in range(3):
            dst.copy_(src, non_blocking=True)
        torch.cuda.synchronize(0)
        e0.record()
        for _ in range(20):
            dst.copy_(src, non_blocking=True)
        e1.record()
        torch.cuda.synchronize(0)
        t = e0.elapsed_time(e1) / 20
        print(f"{mb:3d} MiB copy engine GPU0 -> GPU1: {t * 1e3:7.1f} us  {mb * 2**20 / t / 1e6:7.1f} GB/s")
