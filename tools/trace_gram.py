"""Pipeline trace of one Gram launch (experiment build -DFB_TRACE=1): when did each role of a CTA reach each tile?

    python tools/trace_gram.py build        # here: nvcc -> focal_b200/libfocal_b200_trace.so (headline shapes only)
    python tools/trace_gram.py run [stage]  # on the GPU box; stage = temporal (default) | nce_grad | nce_rowsum
Prints, for CTA 0 and a CTA in the middle of the grid, per column tile (clk relative to the CTA's first event):
producer: B-tile load issued; issuer: UMMA#1 issue start / done, UMMA#2 issue start / done; epilogue warpgroups: wait start,
tile available (s_full), tile done.
"""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
TILES, ROLES, TAGS, BLOCKS = 96, 6, 4, 160


def build():
    from focal_b200 import build as b
    extra = {k: v for k, v in (kv.split("=") for kv in os.environ.get("FB_DEFS", "").split(",") if kv)}
    print(b.build_variant("trace", dict({"FB_TRACE": 1, "FB_FAST_BUILD": 1}, **extra)))


def run(stage="temporal"):
    import numpy as np
    import torch
    from focal_b200 import _cabi
    lib = _cabi.load(os.path.join(ROOT, "focal_b200", "libfocal_b200_trace.so"))
    B, D, M, S = int(os.environ.get("FB_B", 8192)), int(os.environ.get("FB_D", 256)), 2, 4
    prec = int(os.environ.get("FB_PREC", 0))
    R = int(os.environ.get("FB_R", 1))                 # FB_R = 8: one rank's share of a row-sharded step (owned rows = 1/8)
    cfg = _cabi.FocalCfg(B=B, S=S, M=M, D=D, temperature=0.5, margin=1.0, w_shared=1, w_private=1, w_orth=3, w_rank=5,
                         need_grad=1, terms=7, seq_begin=0, seq_end=B // S // R, precision=prec)
    info = _cabi.FocalWsInfo()
    assert lib.focal_b200_workspace_info(C.byref(cfg), C.byref(info)) == 0
    raw = torch.zeros(info.total_bytes + 1024, dtype=torch.uint8, device="cuda")
    off = (-raw.data_ptr()) % 1024
    wsp, wsn = C.c_void_p(raw.data_ptr() + off), C.c_size_t(info.total_bytes)
    feats = [torch.randn(B, D, device="cuda") for _ in range(2 * M)]
    fptr = _cabi.ptr_array([t.data_ptr() for t in feats])
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    ref = C.byref(cfg)
    for it in range(3):
        assert lib.focal_b200_prologue(ref, fptr, wsp, wsn, st) == 0
        assert lib.focal_b200_nce_rowsum(ref, wsp, wsn, st) == 0
        assert lib.focal_b200_nce_lse(ref, wsp, wsn, 0, st) == 0
        torch.cuda.synchronize()
        lib.focal_b200_debug_trace_clear()
        fn = {"temporal": lib.focal_b200_temporal, "nce_grad": lib.focal_b200_nce_grad,
              "nce_rowsum": lib.focal_b200_nce_rowsum}[stage]
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        assert fn(ref, wsp, wsn, st) == 0
        e1.record()
        torch.cuda.synchronize()
    print(f"{stage}: {e0.elapsed_time(e1) * 1e3:.1f} us")
    n = BLOCKS * ROLES * TILES * TAGS
    buf = (C.c_longlong * n)()
    lib.focal_b200_debug_trace.argtypes = [C.c_void_p, C.c_size_t]
    assert lib.focal_b200_debug_trace(buf, n) == 0
    tr = np.frombuffer(buf, dtype=np.int64).reshape(BLOCKS, ROLES, TILES, TAGS)
    if os.environ.get("FB_TIMELINE"):
        timeline(tr)
        return
    for cta in (0, 74):
        t = tr[cta]
        nz = t[t > 0]
        if nz.size == 0:
            continue
        t0 = nz.min()
        rel = np.where(t > 0, t - t0, -1)
        print(f"--- CTA {cta}: tile | producer load | issuer U1 start/done U2 start/done | epilogue wg: wait-start/avail/done ...")
        for n_ in range(40):
            row = [f"{n_:3d}", f"P {rel[0, n_, 0]:6d}", "I " + " ".join(f"{rel[1, n_, k]:6d}" for k in range(4))]
            for g in range(4):
                if (rel[2 + g, n_] >= 0).any():
                    row.append(f"E{g} " + " ".join(f"{rel[2 + g, n_, k]:6d}" for k in range(3)))
            print(" | ".join(row))
        # per-tile period of the issuer over tiles 8..40
        u1 = rel[1, 8:40, 0]
        print("   mean UMMA#1 issue period (tiles 8-40):", float(np.diff(u1[u1 >= 0]).mean()) if (u1 >= 0).sum() > 2 else None)


def timeline(tr):
    """Whole-kernel timeline of a few CTAs (clk since the CTA entered the kernel): set-up, per piece the A request, per
    tile UMMA#1 issue / UMMA#2 done / slowest epilogue done, O written out, kernel exit."""
    import numpy as np
    for cta in (0, 37, 74, 111, 147):
        t = tr[cta]
        t0 = t[0, 0, 2]
        if t0 <= 0:
            continue
        rel = np.where(t > 0, t - t0, -1)
        ntile = int((rel[1, :, 0] >= 0).sum())
        print(f"--- CTA {cta}: set-up done {rel[0, 0, 3]}, {ntile} tiles, exit {rel[0, 1, 2]} clk")
        for n_ in range(ntile):
            a = rel[0, n_, 1]
            epi = max(int(rel[2 + g, n_, 2]) for g in range(4))
            odone = max(int(rel[2 + g, n_, 3]) for g in range(4))
            print(f"   tile {n_:3d}: " + (f"A req {a:7d} " if a >= 0 else " " * 14) + f"B req {rel[0, n_, 0]:7d}  U1 {rel[1, n_, 0]:7d}  U2 done {rel[1, n_, 3]:7d}"
                  f"  epi done {epi:7d}" + (f"  O out {odone:7d}" if odone >= 0 else ""))


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "build":
        build()
    else:
        run(*(sys.argv[2:3] if len(sys.argv) > 2 else []))
