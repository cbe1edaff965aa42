#!/usr/bin/env python
"""Benchmark of the FOCAL contrastive-loss hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

A "step" is one pass of the hot path -- ``loss = FOCALLoss(f1, f2); loss.backward()`` w.r.t. all 2M feature
tensors -- over one batch of synthetic MOD-shaped embeddings: B = 8192 rows (2048 sequences of S = 4 windows),
M = 2 modalities, D = 256 (128 shared + 128 private), T = 0.5.  For N > 1 the same global batch is row-sharded
over the ranks (strong scaling; launched by torchrun, one rank per GPU; the ranks exchange operands through NVLink peer
memory, NCCL carries only the set-up handshake and the timing reductions).

``value`` is device-resident throughput (inputs in HBM, CUDA-graph replays, CUDA events, max over ranks); ``e2e`` is the
same step through ``FOCALLoss.forward/backward`` with pinned HOST inputs: the H2D copy of every step and the D2H read of
every step's loss are inside the timed region (copy of step k+1 and read-back of step k-1 overlap the kernels of step k).

``--impl reference`` times the reference's CPU implementation of the path on the host cores: the UNMODIFIED reference
``FOCALLoss`` from ``oracle/_ref`` (placed there by ``python -m oracle.build_ref``, travels to the GPU box like a built
``.so``; kind = "reference"), or -- only when that is absent -- its op-for-op port ``oracle/focal_ref_port.py``
(kind = "port"), on a bounded sample of the workload.

Roofline fractions are quoted against the measured peak that matches the clock sampled DURING the timed region: the burst
figure of ``MEASURED_PEAKS.json`` when the SM clock stayed near its maximum (short runs), the sustained figure when
the sampler saw the power-capped clock.  ``sustained`` is a second, >= 2 s timed pass with >= 50 clock samples.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "FOCAL loss fwd+bwd samples/s @B=8192"
UNIT = "samples/s"
WORKLOADS = {
    # BASELINE.json metric configuration (configs[1] dims at the headline batch): the default and the only bench line
    "headline": dict(B=8192, S=4, M=2, D=256, T=0.5, mods=("seismic", "audio"), terms=7),
    # configs[0]: the reference's own CPU-runnable case
    "cfg1": dict(B=128, S=4, M=2, D=128, T=0.5, mods=("seismic", "audio"), terms=7),
    # the other BASELINE.json configs, for profiles/ (python bench.py --workload cfg3 ...)
    "cfg2": dict(B=1024, S=4, M=2, D=256, T=0.5, mods=("seismic", "audio"), terms=7),
    "cfg3": dict(B=4096, S=4, M=3, D=256, T=0.07, mods=("acc", "gyr", "mag"), terms=7),
    "cfg4": dict(B=65536, S=1, M=2, D=128, T=0.5, mods=("seismic", "audio"), terms=1),      # global InfoNCE only
    "cfg5": dict(B=16384, S=4, M=4, D=256, T=0.5, mods=("m0", "m1", "m2", "m3"), terms=7),
    "cfg5w": dict(B=16384, S=4, M=4, D=512, T=0.5, mods=("m0", "m1", "m2", "m3"), terms=7),  # shared 256 + private 256
}
for _w in WORKLOADS.values():
    _w.update(margin=1.0, weights=(1.0, 1.0, 3.0, 5.0))
WORKLOAD = WORKLOADS["headline"]
L2_BYTES = 126 * 2 ** 20


def f_alg(B, M, D, S, terms=7):
    """Algorithmic FLOPs of one step, SURVEY.md §8d: 12 B^2 D (M^2/S + M)  (InfoNCE part + temporal part)."""
    nce = 12.0 * M * M * B * B * D / S if terms & 1 else 0.0
    tmp = 12.0 * M * B * B * D if terms & 4 else 0.0
    return nce + tmp


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as fh:
            p = json.load(fh)
        cl = p.get("clocks_under_load", {})
        return dict(tflops_sustained=float(p.get("bf16_tflops_sustained", p.get("bf16_tflops", 1413.7))),
                    tflops_burst=float(p.get("bf16_tflops", 1662.8)), hbm=float(p.get("hbm_gbs", 6541.1)),
                    sustained_mhz=float(cl.get("sm_mhz_median", 1350.0)), max_mhz=float(p.get("sm_max_mhz", 1965.0)),
                    source="measured")
    # /opt/skills/guides/B200_PROFILING.md fallback: 1.59 PFLOP/s burst, ~1.4 sustained, 6.65 TB/s
    return dict(tflops_sustained=1400.0, tflops_burst=1590.0, hbm=6650.0, sustained_mhz=1300.0, max_mhz=1965.0,
                source="fallback")


def tensor_peak_for(peaks, clocks, tf32=False):
    """The measured dense peak that matches the clock seen DURING the timed region: burst when the SM clock stayed within
    10 % of its maximum, sustained (the power-capped figure) otherwise.  TF32 runs at half the bf16 rate."""
    mhz = (clocks or {}).get("sm_mhz")
    burst = mhz is None or mhz >= 0.9 * peaks["max_mhz"]
    val = peaks["tflops_burst"] if burst else peaks["tflops_sustained"]
    if tf32:
        val *= 0.5
    kind = ("burst" if burst else "sustained") + (" bf16 x 0.5 (tf32)" if tf32 else " bf16")
    return val, f"{peaks['source']} ({kind}; sampled SM clock {mhz} MHz)"


# -------------------------------------------------------------------------------------------------
# clocks sampler (nvidia-smi during the timed region)
# -------------------------------------------------------------------------------------------------
class ClockSampler:
    """Samples SM clock, power and throttle reasons of one GPU during the timed region (NVML; nvidia-smi fallback)."""

    def __init__(self, index: int):
        self.index = index
        self.samples = []          # (sm_mhz, sm_max_mhz, power_w, reasons bitmask)
        self._stop = threading.Event()
        self._t = threading.Thread(target=self._run, daemon=True)
        self._nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self._nvml = pynvml
            self._h = pynvml.nvmlDeviceGetHandleByIndex(self._physical_index(index))
            self._max = pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self._nvml = None

    @staticmethod
    def _physical_index(index: int) -> int:
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        if vis:
            ids = [v.strip() for v in vis.split(",") if v.strip()]
            if index < len(ids) and ids[index].isdigit():
                return int(ids[index])
        return index

    def _sample_nvml(self):
        n = self._nvml
        sm = n.nvmlDeviceGetClockInfo(self._h, n.NVML_CLOCK_SM)
        try:
            pw = n.nvmlDeviceGetPowerUsage(self._h) / 1000.0
        except Exception:
            pw = float("nan")
        try:
            rs = n.nvmlDeviceGetCurrentClocksEventReasons(self._h)
        except Exception:
            rs = n.nvmlDeviceGetCurrentClocksThrottleReasons(self._h)
        self.samples.append((float(sm), float(self._max), pw, int(rs)))

    def _sample_smi(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        out = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i",
                              str(self.index)], capture_output=True, text=True, timeout=5).stdout
        parts = [x.strip() for x in out.strip().split(",")]
        if len(parts) >= 7:
            bits = 0
            for k, bit in enumerate((0x8, 0x40, 0x20, 0x4)):
                if parts[3 + k].lower().startswith("active"):
                    bits |= bit
            self.samples.append((float(parts[0]), float(parts[1]), float(parts[2]), bits))

    def _run(self):
        while not self._stop.is_set():
            try:
                if self._nvml is not None:
                    self._sample_nvml()
                else:
                    self._sample_smi()
            except Exception:
                pass
            self._stop.wait(0.02 if self._nvml is not None else 0.2)

    def __enter__(self):
        self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._t.join(timeout=10)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        # NVML clocks-event-reason bits: 0x4 sw_power_cap, 0x8 hw_slowdown, 0x20 sw_thermal, 0x40 hw_thermal
        names = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}
        allbits = 0
        for s_ in self.samples:
            allbits |= s_[3]
        return {"sm_mhz": statistics.median(s_[0] for s_ in self.samples),
                "sm_min_mhz": min(s_[0] for s_ in self.samples),
                "sm_max_mhz": max(s_[1] for s_ in self.samples),
                "power_w_max": max(s_[2] for s_ in self.samples),
                "reasons": [n for b, n in names.items() if allbits & b], "samples": len(self.samples),
                "source": "nvml" if self._nvml is not None else "nvidia-smi"}


# -------------------------------------------------------------------------------------------------
# CPU baseline (oracle port of the reference) -- bounded sample of the workload
# -------------------------------------------------------------------------------------------------
def _reference_callable(w):
    """(kind, fn(f1, f2) -> loss): the unmodified reference module from oracle/_ref, else the op-for-op port."""
    import types

    from oracle.focal_oracle import FocalConfig
    try:
        from oracle.build_ref import import_reference_loss, ref_available
        if ref_available():
            cls = import_reference_loss()
            args_ns = types.SimpleNamespace(
                device="cpu", model="DeepSense", tag=None,
                dataset_config={"modality_names": list(w["mods"]), "seq_len": w["S"],
                                "FOCAL": {"temperature": {"DeepSense": w["T"], "SW_Transformer": 0.07},
                                          "inter_rank_margin": w["margin"],
                                          "shared_contrastive_loss_weight": w["weights"][0],
                                          "private_contrastive_loss_weight": w["weights"][1],
                                          "orthogonal_loss_weight": w["weights"][2],
                                          "rank_loss_weight": w["weights"][3]}})
            mod = cls(args_ns)
            return "reference", (lambda f1, f2: mod(f1, f2))
    except Exception as exc:                                  # noqa: BLE001 -- fall back to the port, but say so
        print(f"bench.py: oracle/_ref unusable ({exc!r}); timing the port", file=sys.stderr)
    from oracle.focal_ref_port import focal_loss_port
    cfg = FocalConfig(modalities=list(w["mods"]), seq_len=w["S"], temperature=w["T"], margin=w["margin"])
    return "port", (lambda f1, f2: focal_loss_port(f1, f2, cfg))


def cpu_reference_time(B_sample: int, steps: int, warmup: int, w=None):
    """Seconds per fwd+bwd of the reference's CPU implementation on B_sample rows of workload w.  Returns (kind, times)."""
    from oracle.focal_oracle import make_iid
    w = w or WORKLOAD
    kind, fn = _reference_callable(w)
    f1, f2 = make_iid(0, w["mods"], B_sample, w["D"])
    f1 = {m: v.requires_grad_(True) for m, v in f1.items()}
    f2 = {m: v.requires_grad_(True) for m, v in f2.items()}
    times = []
    for it in range(warmup + steps):
        for v in list(f1.values()) + list(f2.values()):
            v.grad = None
        t0 = time.perf_counter()
        loss = fn(f1, f2)
        loss.backward()
        float(loss.detach())
        t1 = time.perf_counter()
        if it >= warmup:
            times.append(t1 - t0)
    return kind, times


def pick_cpu_sample(budget_s: float, steps: int, warmup: int) -> int:
    """Largest B in {128, 256, 512} (capped at the workload's batch) whose (steps + warmup) run fits the budget; the
    reference's cost grows ~ B^2."""
    cap = WORKLOAD["B"]
    t128 = min(cpu_reference_time(min(128, cap), 2, 1)[1])
    best = min(128, cap)
    for B in (256, 512):
        if B <= cap and (steps + warmup) * t128 * (B / 128) ** 2 * 1.3 < budget_s:
            best = B
    return best


def cfg1_reference(iters=30, warmup=5):
    """the reference at configs[0] (B=128, M=2, S=4, D=128), 5 warm-up + 30 timed, median."""
    w = WORKLOADS["cfg1"]
    kind, ts = cpu_reference_time(w["B"], iters, warmup, w)
    med = statistics.median(ts)
    return {"workload": "cfg1: B=128, M=2, S=4, D=128, T=0.5 (BASELINE.json configs[0])", "kind": kind,
            "ms_median": med * 1e3, "samples_per_s": w["B"] / med, "iters": iters, "warmup": warmup}


def cpu_baseline_record(budget_s, steps, warmup):
    import torch
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    w = WORKLOAD
    Bs = pick_cpu_sample(budget_s, steps, warmup)
    kind, ts = cpu_reference_time(Bs, steps, warmup)
    what = ("unmodified reference FOCALLoss (oracle/_ref, /root/reference/src/models/loss.py)" if kind == "reference"
            else "op-for-op port of the reference FOCALLoss (oracle/focal_ref_port.py)")
    sample = (f"{what}, fwd+bwd on CPU, B={Bs} rows of the B={w['B']} workload (D={w['D']}, M={w['M']}, S={w['S']}); "
              f"cost grows ~B^2 (the reference materialises [S,2b,2b,d]: 34 GB per call at B=8192, it cannot run there)")
    rec = {"value": Bs / statistics.mean(ts), "unit": UNIT, "cores": cores, "kind": kind, "sample": sample,
           "sample_batch": Bs, "ms_per_step": statistics.mean(ts) * 1e3}
    return rec


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    rec = cpu_baseline_record(150.0, args.steps, args.warmup)
    try:
        rec["cfg1"] = cfg1_reference()
    except Exception as exc:                                  # noqa: BLE001
        rec["cfg1"] = {"error": repr(exc)}
    w = WORKLOAD
    line = {
        "impl": "reference", "metric": METRIC, "value": rec["value"], "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": rec["ms_per_step"], "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(), "global_batch": w["B"], "cpu_sample_batch": rec["sample_batch"]},
        "cpu_baseline": rec,
        "e2e": {"value": rec["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


def workload_name():
    w = WORKLOAD
    return (f"FOCAL loss fwd+bwd, synthetic MOD-shaped embeddings: B={w['B']} (b={w['B'] // w['S']} sequences x S={w['S']}), "
            f"M={w['M']} modalities, D={w['D']} (shared {w['D'] // 2} + private {w['D'] // 2}), T={w['T']}, "
            f"weights {w['weights']}")


# -------------------------------------------------------------------------------------------------
# our arm
# -------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist

    import focal_b200
    from focal_b200 import _cabi
    from focal_b200.engine import FocalEngine, FocalHyper, resolve_precision

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    group = None
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        group = dist.group.WORLD
    if args.gpus != world and rank == 0:
        print(f"bench.py: --gpus {args.gpus} but WORLD_SIZE={world}; using WORLD_SIZE", file=sys.stderr)

    w = WORKLOAD
    B, S, M, D = w["B"], w["S"], w["M"], w["D"]
    mods = list(w["mods"])
    if B % (S * world):
        raise SystemExit("global batch does not shard over the ranks")
    Bl = B // world
    hp = FocalHyper(tuple(mods), S, w["T"], w["margin"], *w["weights"], False, w["terms"], args.precision)
    engine = FocalEngine(hp, process_group=group)
    fp32_mode = resolve_precision(hp, B, D) == _cabi.FOCAL_PREC_FP32

    # synthetic inputs: NSETS different batches so that consecutive steps read their inputs from HBM, not L2
    nsets = min(16, max(2, math.ceil(2 * L2_BYTES / (2 * M * B * D * 4))))
    gen = torch.Generator(device="cpu").manual_seed(1234)
    host_sets, dev_sets = [], []
    for s_ in range(nsets):
        full = [torch.randn(B, D, generator=gen, dtype=torch.float32) for _ in range(2 * M)]
        mine = [t[rank * Bl:(rank + 1) * Bl].contiguous().pin_memory() for t in full]
        host_sets.append(mine)
        dev_sets.append([t.to(dev) for t in mine])

    def as_dicts(tensors):
        return ({m: tensors[i] for i, m in enumerate(mods)}, {m: tensors[M + i] for i, m in enumerate(mods)})

    def step(k):
        f1, f2 = as_dicts(dev_sets[k % nsets])
        return engine.loss_and_grads(f1, f2, True)

    def sync_all():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- warm-up: W >= 3 steps (the engine runs the first one eagerly, captures the launch sequence on the second and
    # replays it from then on; the captured graph serves every input set -- pointers travel through the workspace table)
    n_warm = max(args.warmup, 3)
    # (the sampler is built BEFORE the warm-up: nvmlInit takes tens of milliseconds, and that much idle time between the
    # warm-up and the timed steps lets the GPU drop out of its boost state -- short timed regions then start on ramping
    # clocks: 20-step runs measured 0.514 ms/step against 0.495 for 100 steps)
    sampler = ClockSampler(local_rank) if rank == 0 else None
    for k in range(n_warm):
        loss5, grads = step(k)          # held across the next step exactly like in the timed loop (allocator steady state)
    sync_all()

    # ---- timed region: exactly K steps, device-timed, max over ranks
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if sampler:
        sampler.__enter__()
    sync_all()
    ev0.record()
    h0 = time.perf_counter()
    for k in range(args.steps):
        loss5, grads = step(k)
    host_ms = (time.perf_counter() - h0) * 1e3 / args.steps     # time the host needs to enqueue one step
    ev1.record()
    sync_all()
    if sampler:
        sampler.__exit__()
    ms_total = ev0.elapsed_time(ev1)
    t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step = float(t.item()) / args.steps
    value = B / (ms_step * 1e-3)
    clocks = sampler.summary() if sampler else None

    # ---- sustained pass: >= 2 s of back-to-back steps with the clocks sampled throughout (what the step costs once the
    # power cap has pulled the SM clock down); reported beside the K-step figure, never instead of it
    sustained = None
    if world == 1 and args.sustain_seconds > 0:
        n_sus = max(50, int(1.3 * args.sustain_seconds * 1e3 / max(ms_step, 1e-3)) + 1)
        with ClockSampler(local_rank) as s2:
            torch.cuda.synchronize()
            ev0.record()
            for k in range(n_sus):
                step(k)
            ev1.record()
            torch.cuda.synchronize()
        sus_ms = ev0.elapsed_time(ev1) / n_sus
        sustained = {"steps": n_sus, "seconds": sus_ms * n_sus * 1e-3, "ms_per_step": sus_ms,
                     "value": B / (sus_ms * 1e-3), "unit": UNIT, "clocks": s2.summary()}

    # ---- end-to-end: pinned host inputs -> H2D -> loss + grads -> D2H of the loss, through the module API
    args_ns = make_args(mods, S, w, group, args.precision)
    module = focal_b200.FOCALLoss(args_ns).to(dev)
    if w["terms"] != 7:
        module._engine = FocalEngine(hp, process_group=group)      # sub-set of the terms (cfg4: InfoNCE only)
    # Three device staging sets fed from pinned host memory on a copy stream, like a pinned-memory data loader with
    # non_blocking copies: the H2D copy of step k+1 overlaps the kernels of step k.  Every step copies its own inputs
    # host->device and reads its own loss back to the host; the read-back is asynchronous (pinned buffer + event) and the
    # host picks the value of step k up after it has enqueued step k+1, so the GPU never idles on the host round trip.
    NSLOT = 3
    stages_dev = [[torch.empty(Bl, D, device=dev) for _ in range(2 * M)] for _ in range(NSLOT)]
    copy_stream = torch.cuda.Stream(device=dev)
    copied = [torch.cuda.Event() for _ in range(NSLOT)]
    consumed = [torch.cuda.Event() for _ in range(NSLOT)]
    landed = [torch.cuda.Event() for _ in range(NSLOT)]
    host_loss = [torch.empty((), dtype=torch.float32).pin_memory() for _ in range(NSLOT)]
    e2e_losses = []

    def prefetch(k):
        slot = k % NSLOT
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(consumed[slot])            # the step that last used this slot has finished with it
            for dst, src in zip(stages_dev[slot], host_sets[k % nsets]):
                dst.copy_(src, non_blocking=True)
            copied[slot].record(copy_stream)

    def collect(k):
        slot = k % NSLOT
        landed[slot].synchronize()                            # the loss of step k is in host memory
        e2e_losses.append(float(host_loss[slot]))

    def e2e_step(k, collect_prev=True):
        slot = k % NSLOT
        prefetch(k + 1)                                       # next step's inputs travel while this step computes
        torch.cuda.current_stream().wait_event(copied[slot])
        xs = [s_.detach().requires_grad_(True) for s_ in stages_dev[slot]]
        f1, f2 = as_dicts(xs)
        loss = module(f1, f2)
        loss.backward()
        consumed[slot].record()
        host_loss[slot].copy_(loss.detach(), non_blocking=True)   # D2H read of the step's result
        landed[slot].record()
        if collect_prev:
            collect(k - 1)                                    # ... picked up one step later

    for ev in consumed:
        ev.record()
    prefetch(0)
    n_e2e_warm = max(args.warmup, 3)
    for k in range(n_e2e_warm):
        e2e_step(k, collect_prev=k > 0)
    collect(n_e2e_warm - 1)
    sync_all()
    ev0.record()
    for k in range(n_e2e_warm, n_e2e_warm + args.steps):
        e2e_step(k, collect_prev=k > n_e2e_warm)
    collect(n_e2e_warm + args.steps - 1)                      # the last loss is read inside the timed region too
    ev1.record()
    sync_all()
    te = torch.tensor([ev0.elapsed_time(ev1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_ms = float(te.item()) / args.steps
    e2e = {"value": B / (e2e_ms * 1e-3), "unit": UNIT, "ms_per_step": e2e_ms,
           "h2d_bytes_per_step": 2 * M * Bl * D * 4 * world, "d2h_bytes_per_step": 4 * world}

    # ---- per-stage device times (separate pass; events between the C-ABI stages on the launching stream)
    stages = stage_breakdown(engine, dev_sets, mods, M, args.steps, world, group) if world == 1 else None

    out = None
    if rank == 0:
        peaks = measured_peaks()
        F = f_alg(B, M, D, S, w["terms"])
        peak_tf, peak_src = tensor_peak_for(peaks, clocks, tf32=False)
        roof = None
        if stages is not None:
            # dominant launch: the fused temporal distance/ranking pass (2M calls, fwd+bwd in one launch); for the
            # InfoNCE-only workload the backward Gram pass (2 of the 3 GEMM-equivalents of the InfoNCE term)
            if w["terms"] & 4:
                kname, F_k, t_k = "gram_kernel<TMP_BWD> (temporal ranking fwd+bwd, fused)", 12.0 * M * B * B * D, stages["temporal"]
            else:
                kname, F_k, t_k = "gram_kernel<NCE_BWD> (InfoNCE backward)", 8.0 * M * M * B * B * D / S, stages["nce_grad"]
            traffic = ncu_tensor = None
            for tname in ("r2_traffic.json", "r1_traffic.json"):
                tpath = os.path.join(ROOT, "profiles", tname)
                if args.workload == "headline" and not fp32_mode and os.path.exists(tpath):
                    with open(tpath) as fh:
                        tj = json.load(fh)
                    traffic = tj.get("gram_kernel_tmp_bwd_dram_bytes_per_launch")
                    ncu_tensor = tj.get("gram_kernel_tmp_bwd_tensor_pipe_active_pct_elapsed")
                    break
            t_s = max(t_k, 1e-6) * 1e-3
            # SURVEY 8d counts 3 GEMM-equivalents per contraction; the fused temporal pass executes 2 of them (the
            # symmetric column-side gradient comes for free), the InfoNCE backward pass executes what it counts.
            # fp32 mode: every executed GEMM is three bf16 passes (hi*hi + hi*lo + lo*hi).
            executed = ((2.0 / 3.0) if (w["terms"] & 4) else 1.0) * (3.0 if fp32_mode else 1.0)
            roof = {"bound": "tensor", "kernel": kname, "achieved": F_k / t_s / 1e12, "peak": peak_tf,
                    "unit": "TFLOP/s", "frac": F_k / t_s / 1e12 / peak_tf,
                    "peak_source": peak_src, "traffic": traffic,
                    "alg_flops_per_launch": F_k, "launch_ms": t_k,
                    "executed_frac": executed * F_k / t_s / 1e12 / peak_tf,
                    "frac_vs_sustained_peak": F_k / t_s / 1e12 / peaks["tflops_sustained"],
                    "frac_vs_burst_peak": F_k / t_s / 1e12 / peaks["tflops_burst"],
                    "ncu_tensor_pipe_active_pct": ncu_tensor,
                    "note": "achieved = ALGORITHMIC FLOPs (SURVEY 8d: 3 GEMM-equivalents per contraction) / measured "
                            "launch time; frac = achieved / the measured dense bf16 peak that matches the SM clock sampled "
                            "during the timed region (burst at ~max clock, sustained under the power cap); executed_frac "
                            "counts the tensor-core FLOPs the kernel really issues; ncu's sm__pipe_tensor_cycles_active "
                            "(profiles/) is quoted beside it"}
        row_kernels = None
        if stages is not None:
            # the O(B D) kernels against the HBM roofline: algorithmic bytes = what must cross HBM once
            opb = 4.0 if fp32_mode else 2.0                                    # operand bytes per element (hi + lo images)
            feat_bytes = 2.0 * M * B * D * 4                                   # 2M fp32 [B, D] tensors
            pro_bytes = feat_bytes + 2 * (2.0 * M * B * D * opb)               # read features, write both operand sets
            fin_bytes = feat_bytes + feat_bytes                                # read features, write gradients
            if w["terms"] & 1:
                fin_bytes += 2.0 * M * B * D * 4                               # + the InfoNCE accumulators (dz)
            if w["terms"] & 4:
                fin_bytes += 2.0 * M * B * D * 4                               # + the temporal accumulators (dx)
            row_kernels = {"peak_GBps": peaks["hbm"], "peak_source": peaks["source"]}
            for name, nbytes, key in (("prologue", pro_bytes, "prologue"), ("finalize", fin_bytes, "finalize")):
                t_s = max(stages[key], 1e-6) * 1e-3
                row_kernels[name] = {"alg_bytes": nbytes, "ms": stages[key], "GBps": nbytes / t_s / 1e9,
                                     "frac_hbm": nbytes / t_s / 1e9 / peaks["hbm"]}
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            cpu = cpu_baseline_record(12.0, 3, 1)
            try:
                cpu["cfg1"] = cfg1_reference(iters=30, warmup=5)
            except Exception as exc:                              # noqa: BLE001
                cpu["cfg1"] = {"error": repr(exc)}
        sus_frac = None
        if sustained is not None:
            pk, src = tensor_peak_for(peaks, sustained["clocks"])
            sustained["tensor_roofline_frac"] = F / (sustained["ms_per_step"] * 1e-3) / 1e12 / pk
            sustained["peak"] = pk
            sustained["peak_source"] = src
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": n_warm, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "bf16x3 (fp32 mode)" if fp32_mode else "bf16", "data": "synthetic",
            "config": {"workload": workload_name(), "global_batch": B, "rows_per_gpu": Bl,
                       "parallelism": (f"row-sharded x{world}, "
                                       + ("exchange by kernel stores into NVLink peer memory + device barriers"
                                          if any(v is not None for v in getattr(engine.backend, "_peers", {}).values())
                                          else "NCCL all-gather(features, row sums) + all-reduce(loss)"))
                       if world > 1 else "single GPU",
                       "l2": f"inputs rotate over {nsets} batches ({nsets * 2 * M * B * D * 4 / 2 ** 20:.0f} MiB"
                             + (" > 126 MiB L2)" if nsets * 2 * M * B * D * 4 > L2_BYTES else ", fits L2: small side workload)"),
                       "tiles": ("split-bf16 operands (hi + lo, 16 significant bits), three tcgen05 kind::f16 passes per "
                                 "product, fp32 accumulation" if fp32_mode
                                 else "bf16 operands, fp32 accumulation (tcgen05 kind::f16)")},
            "tensor_roofline_frac": F / (ms_step * 1e-3) / 1e12 / (peak_tf * world),   # of the N GPUs' peak
            "tensor_roofline_peak": {"tflops_per_gpu": peak_tf, "source": peak_src},
            "alg_tflops": F / (ms_step * 1e-3) / 1e12, "sustained": sustained,
            "roofline": roof, "row_kernels_hbm": row_kernels, "cpu_baseline": cpu, "e2e": e2e,
            "gpu_launches": args.steps * launches_per_step(engine, B, D, world),
            "stages_ms": stages, "host_enqueue_ms_per_step": host_ms,
            "cuda_graph_replays": engine.graph_replays, "cuda_graph_captures": engine.graph_captures,
            "clocks": clocks,
            "loss": float(loss5[0].item()),
        }
        print(json.dumps(out))
    sys.stdout.flush()
    if world > 1:
        # Tear down in a fixed order (graphs hold captured NCCL work) and never let a teardown problem turn into a hung
        # job: a watchdog hard-exits the rank if the orderly shutdown has not finished after 30 s.
        def _bail():
            os._exit(0)
        wd = threading.Timer(30.0, _bail)
        wd.daemon = True
        wd.start()
        engine._steps.clear()
        if module._engine is not None:
            module._engine._steps.clear()
        torch.cuda.synchronize()
        dist.barrier()
        engine.backend.close()
        if module._engine is not None and module._engine.backend is not engine.backend:
            module._engine.backend.close()
        dist.barrier()
        dist.destroy_process_group()
        wd.cancel()
    return 0


def launches_per_step(engine, B, D, world):
    """Kernels of libfocal_b200.so per step, counted from the plan of this workload (plan.h / focal_b200.cu).
    Third-generation row kernels (S in {1, 2, 4}, D/2 * S a multiple of 128 and <= 512, D <= 256, <= 8 tensors):
    prologue_v3 (zeroes the padding rows itself), nce_rowsum, nce_lse, nce_grad, temporal, finalize_v3 (its last block is
    the loss reduction).  Older row kernels: [zero_pad], prologue, [intra], ..., finalize, loss_reduce.  The collective
    multi-GPU path adds nce_lse over all rows; the peer path has no extra launches (its barriers live at the head of
    the consuming kernels).  + set_ptrs before every graph replay."""
    import ctypes as C
    import os

    from focal_b200 import _cabi
    hp = engine.hp
    be = engine.backend
    cfg = be._cfg(hp, B, D, True, (0, B // hp.seq_len))
    info = _cabi.FocalWsInfo()
    _cabi.check(be.lib.focal_b200_workspace_info(C.byref(cfg), C.byref(info)), "workspace_info")
    peer = world > 1 and any(v is not None for v in getattr(be, "_peers", {}).values())
    S, d, nT = hp.seq_len, D // 2, 2 * len(hp.modalities)
    v3 = (not hp.no_private and D % 2 == 0 and S in (1, 2, 4) and (d * S) % 128 == 0 and d * S // 128 <= 4 and D <= 256
          and nT <= 8 and os.environ.get("FOCAL_B200_ROW_KERNELS") not in ("v1", "v2"))
    n = 0
    if not peer and not v3:
        n += 1 if (info.bpad != info.b or info.Bpad != B) else 0    # zero_pad_kernel (peer workspaces start zeroed)
    n += 1                                                          # prologue (m_II fused for S in {2, 4})
    n += 0 if S in (2, 4) or v3 or not (hp.terms & 4) else 1        # separate intra_kernel otherwise
    if hp.terms & 1:
        n += 3                                                      # nce_rowsum, nce_lse, nce_grad
        n += 1 if (world > 1 and not peer) else 0                   # collective path: nce_lse again over all rows
    if (hp.terms & 4) and S > 1:
        n += 1                                                      # temporal
    n += 1 if v3 else 2                                             # finalize (+ loss_reduce on the older path)
    n += 1 if engine.use_cuda_graph else 0                          # set_ptrs before every graph replay
    return n


def make_args(mods, S, w, group, precision="auto"):
    import types
    return types.SimpleNamespace(
        device="cuda", model="DeepSense", tag=None, focal_process_group=group, focal_precision=precision,
        dataset_config={"modality_names": list(mods), "seq_len": S,
                        "FOCAL": {"temperature": {"DeepSense": w["T"], "SW_Transformer": 0.07},
                                  "inter_rank_margin": w["margin"],
                                  "shared_contrastive_loss_weight": w["weights"][0],
                                  "private_contrastive_loss_weight": w["weights"][1],
                                  "orthogonal_loss_weight": w["weights"][2], "rank_loss_weight": w["weights"][3]}})


def stage_breakdown(engine, dev_sets, mods, M, steps, world, group):
    """Mean device time of each C-ABI stage over `steps` steps (events on the launching stream, launches pre-queued
    behind a spin kernel so that host launch latency is not part of the figures)."""
    import ctypes as C

    import torch

    from focal_b200 import _cabi
    be = engine.backend
    hp = engine.hp
    names = ["prologue", "nce_rowsum", "nce_lse", "nce_grad", "temporal", "finalize"]
    acc = {n: 0.0 for n in names}
    lib = be.lib
    for k in range(steps):
        feats = dev_sets[k % len(dev_sets)]
        B, D = feats[0].shape
        cfg = be._cfg(hp, B, D, True, (0, B // hp.seq_len))
        ws, info = be.workspace(cfg, feats[0].device)
        stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        wsp, wsn, ref = C.c_void_p(ws.data_ptr()), C.c_size_t(ws.numel()), C.byref(cfg)
        fptr = _cabi.ptr_array([t.data_ptr() for t in feats])
        loss5 = torch.empty(5, device=feats[0].device)
        grads = [torch.empty_like(t) for t in feats]
        gptr = _cabi.ptr_array([g.data_ptr() for g in grads])
        evs = [torch.cuda.Event(enable_timing=True) for _ in range(len(names) + 1)]
        # keep the GPU busy (~0.5 ms) while the host enqueues the whole sequence: the stages then run back to back and
        # the events measure kernels, not the host's launch latency (which exceeds the short row kernels)
        torch.cuda._sleep(1_000_000)
        evs[0].record()
        _cabi.check(lib.focal_b200_prologue(ref, fptr, wsp, wsn, stream), "prologue"); evs[1].record()
        _cabi.check(lib.focal_b200_nce_rowsum(ref, wsp, wsn, stream), "nce_rowsum"); evs[2].record()
        _cabi.check(lib.focal_b200_nce_lse(ref, wsp, wsn, 0, stream), "nce_lse"); evs[3].record()
        _cabi.check(lib.focal_b200_nce_grad(ref, wsp, wsn, stream), "nce_grad"); evs[4].record()
        _cabi.check(lib.focal_b200_temporal(ref, wsp, wsn, stream), "temporal"); evs[5].record()
        _cabi.check(lib.focal_b200_finalize(ref, fptr, wsp, wsn, C.c_void_p(loss5.data_ptr()), gptr, stream),
                    "finalize"); evs[6].record()
        torch.cuda.synchronize()
        for i, n in enumerate(names):
            acc[n] += evs[i].elapsed_time(evs[i + 1])
    return {n: v / steps for n, v in acc.items()}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--precision", choices=["bf16", "fp32", "auto"], default="bf16",
                    help="tile precision of the Gram kernels (north_star's bf16 / fp32 modes)")
    ap.add_argument("--sustain-seconds", type=float, default=2.0,
                    help="length of the second, clock-sampled timed pass (0 disables it)")
    ap.add_argument("--workload", choices=sorted(WORKLOADS), default="headline",
                    help="other BASELINE.json configurations (for profiles/); the bench line is the default")
    args = ap.parse_args()
    global WORKLOAD, METRIC
    WORKLOAD = WORKLOADS[args.workload]
    if args.workload != "headline":
        METRIC = f"FOCAL loss fwd+bwd samples/s @B={WORKLOAD['B']} ({args.workload})"
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
