"""Probe (torchrun, >= 2 GPUs): can torch's symmetric memory give this job peer pointers AND an NVSwitch multicast address?"""
import os
import torch
import torch.distributed as dist
import torch.distributed._symmetric_memory as symm

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
torch.cuda.set_device(dev)
dist.init_process_group("nccl", device_id=dev)
try:
    print(rank, "has_multicast_support:", symm._SymmetricMemory.has_multicast_support(symm.DeviceType.CUDA, dev.index) if hasattr(symm._SymmetricMemory, "has_multicast_support") else "n/a", flush=True)
except Exception as e:
    print(rank, "has_multicast_support error", repr(e), flush=True)
try:
    t = symm.empty(64 * 1024 * 1024, dtype=torch.uint8, device=dev)
    h = symm.rendezvous(t, dist.group.WORLD.group_name)
    print(rank, "rendezvous ok; buffer_ptrs", [hex(p) for p in h.buffer_ptrs], "multicast_ptr", hex(h.multicast_ptr), "signal_pad_ptrs", len(h.signal_pad_ptrs), flush=True)
    # data_ptr alignment and equality with own buffer ptr
    print(rank, "own data_ptr", hex(t.data_ptr()), "aligned1024", t.data_ptr() % 1024 == 0, flush=True)
    t.zero_()
    torch.cuda.synchronize()
    dist.barrier()
except Exception as e:
    print(rank, "symm mem failed:", repr(e), flush=True)
dist.barrier()
dist.destroy_process_group()
