"""Experiment harness: build kernel variants with -D knobs (on the CPU box), time the C-ABI stages of each on the GPU.

    python tools/variant_bench.py build          # here (no GPU): nvcc each variant -> focal_b200/libfocal_b200_<tag>.so
    python tools/variant_bench.py run            # on the GPU box: per-stage CUDA-event times of every built variant
"""
import ctypes as C
import glob
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

VARIANTS = json.loads(os.environ.get("FB_VARIANTS", "null")) or {
    "base": {},
    "poly2": {"FB_POLY_PER8": 2},
    "nosk": {"FB_STREAMK_TMP": -1, "FB_STREAMK_NCE": 0},
    "kb4_3stages": {"FB_KB4_NB": 2},
}


def build():
    from concurrent.futures import ThreadPoolExecutor
    from focal_b200 import build as b
    for f in glob.glob(os.path.join(ROOT, "focal_b200", "libfocal_b200_*.so")):
        os.remove(f)
    with ThreadPoolExecutor(max_workers=4) as ex:
        fast = {} if os.environ.get('FB_FULL_BUILD') else {'FB_FAST_BUILD': 1}
        futs = {tag: ex.submit(b.build_variant, tag, dict(d, **fast)) for tag, d in VARIANTS.items()}
        for tag, f in futs.items():
            print(tag, f.result())


def run(B=int(os.environ.get('FB_B', 8192)), D=int(os.environ.get('FB_D', 256)), M=int(os.environ.get('FB_M', 2)), S=4,
        steps=int(os.environ.get('FB_STEPS', 30))):
    import torch
    from focal_b200 import _cabi
    mods = [f"m{i}" for i in range(M)]
    torch.manual_seed(0)
    sets = [[torch.randn(B, D, device="cuda") for _ in range(2 * M)] for _ in range(8)]
    names = ["prologue", "nce_rowsum", "nce_lse", "nce_grad", "temporal", "finalize"]
    for path in sorted(glob.glob(os.path.join(ROOT, "focal_b200", "libfocal_b200_*.so"))):
        tag = os.path.basename(path)[len("libfocal_b200_"):-3]
        lib = _cabi.load(path)
        def mkcfg(r, R):
            b = B // S
            return _cabi.FocalCfg(B=B, S=S, M=M, D=D, temperature=0.5, margin=1.0, w_shared=1, w_private=1, w_orth=3,
                                  w_rank=5, need_grad=int(os.environ.get('FB_NEED_GRAD', '1')), terms=7,
                                  seq_begin=r * b // R, seq_end=(r + 1) * b // R)
        r, R = (int(v) for v in os.environ.get("FB_SHARD", "0/1").split("/"))     # time the work of rank r of R
        cfg = mkcfg(r, R)
        info = _cabi.FocalWsInfo()
        assert lib.focal_b200_workspace_info(C.byref(cfg), C.byref(info)) == 0
        raw = torch.empty(info.total_bytes + 1024, dtype=torch.uint8, device="cuda")
        off = (-raw.data_ptr()) % 1024
        wsp, wsn = C.c_void_p(raw.data_ptr() + off), C.c_size_t(info.total_bytes)
        loss5 = torch.empty(5, device="cuda")
        grads = [torch.empty(B, D, device="cuda") for _ in range(2 * M)]
        gptr = _cabi.ptr_array([g.data_ptr() for g in grads])
        ref = C.byref(cfg)
        # (with FB_SHARD the other ranks' row sums are whatever the workspace holds: times are valid, the loss is not)
        acc = {n: 0.0 for n in names}
        losses = []
        for k in range(steps + 3):
            feats = sets[k % len(sets)]
            fptr = _cabi.ptr_array([t.data_ptr() for t in feats])
            st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(7)]
            torch.cuda._sleep(1_500_000)     # the host enqueues the whole sequence behind this: no launch latency in the stage times
            ev[0].record()
            assert lib.focal_b200_prologue(ref, fptr, wsp, wsn, st) == 0; ev[1].record()
            assert lib.focal_b200_nce_rowsum(ref, wsp, wsn, st) == 0; ev[2].record()
            assert lib.focal_b200_nce_lse(ref, wsp, wsn, 1 if R > 1 else 0, st) == 0; ev[3].record()
            assert lib.focal_b200_nce_grad(ref, wsp, wsn, st) == 0; ev[4].record()
            assert lib.focal_b200_temporal(ref, wsp, wsn, st) == 0; ev[5].record()
            assert lib.focal_b200_finalize(ref, fptr, wsp, wsn, C.c_void_p(loss5.data_ptr()), gptr, st) == 0
            ev[6].record()
            torch.cuda.synchronize()
            if k >= 3:
                for i, n in enumerate(names):
                    acc[n] += ev[i].elapsed_time(ev[i + 1]) / steps
            if k == 0:
                losses.append(float(loss5[0]))
        tot = sum(acc.values())
        print(f"{tag:10s} total {tot*1e3:7.1f} us | " + " ".join(f"{n}={v*1e3:6.1f}" for n, v in acc.items())
              + f" | loss {losses[0]:.5f}", flush=True)


if __name__ == "__main__":
    if sys.argv[1] == "build":
        build()
    else:
        run()
