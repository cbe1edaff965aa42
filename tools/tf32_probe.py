"""GPU bring-up: which shared-memory layouts does tcgen05.mma kind::tf32 accept for an MN-major B operand?

Round 1 found that the SWIZZLE_128B (16-byte atom) tile that serves bf16 as both a K-major and an MN-major operand returns
zeros as an MN-major operand of kind::tf32.  This script tries the candidates in one go (through focal_b200_debug_umma,
built with -DFOCAL_B200_BRINGUP) and prints the error of each against a float64 product:

  K-major  B (UMMA #1)   / MN-major B (UMMA #2), each with
     layout 2: SWIZZLE_128B, 16-byte chunk c of row r stored at c ^ (r & 7)
     layout 1: SWIZZLE_128B_BASE32B, 32-byte chunk c of row r stored at c ^ (r & 3)     (Swizzle<2,5,2>)
     layout 0: no swizzle, 8 x 16 B core matrices
  and A from shared memory or from tensor memory.

    python tools/tf32_probe.py      (on the GPU box)
"""
import ctypes as C
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

TF32 = 0x100


def tf32_round(x):
    i = x.contiguous().view(torch.int32)
    return ((i + 0x1000) & ~0x1FFF).view(torch.float32)


def img_rows(mat, atom):
    """[R, C] fp32, C % 32 == 0 -> [C/32][R][128 B]; atom = 16: 16-byte chunks ^ (r & 7); 32: 32-byte chunks ^ (r & 3);
    0: plain rows."""
    R, Ccols = mat.shape
    kb = Ccols // 32
    if atom == 0:
        return mat.reshape(R, kb, 32).permute(1, 0, 2).contiguous().view(torch.uint8).reshape(-1)
    per = atom // 4                 # elements per chunk
    nch = 32 // per                 # chunks per row
    x = mat.reshape(R, kb, nch, per).permute(1, 0, 2, 3).contiguous()
    r = torch.arange(R, device=mat.device)
    c = torch.arange(nch, device=mat.device)
    src = c[None, :] ^ (r[:, None] & (nch - 1))
    out = torch.gather(x, 2, src[None, :, :, None].expand(kb, R, nch, per))
    return out.contiguous().view(torch.uint8).reshape(-1)


def img_interleaved(mat):
    """[R, C] fp32 -> [R/8][C/4][8 rows][4 elements]: 128-byte core matrices, row-group major."""
    R, Ccols = mat.shape
    return mat.reshape(R // 8, 8, Ccols // 4, 4).permute(0, 2, 1, 3).contiguous().view(torch.uint8).reshape(-1)


def idesc(M, N, a_mn=0, b_mn=0, fmt=2):
    return (1 << 4) | (fmt << 7) | (fmt << 10) | (a_mn << 15) | (b_mn << 16) | ((N >> 3) << 17) | ((M >> 4) << 24)


def run(lib, a_img, b_img, idsc, a_lbo, a_sbo, a_kstep, b_lbo, b_sbo, b_kstep, ksteps, ncols, flags):
    out = torch.full((128, ncols), float("nan"), device="cuda", dtype=torch.float32)
    rc = lib.focal_b200_debug_umma(C.c_void_p(a_img.data_ptr()), a_img.numel(), C.c_void_p(b_img.data_ptr()),
                                   b_img.numel(), idsc, a_lbo, a_sbo, a_kstep, b_lbo, b_sbo, b_kstep, ksteps, ncols,
                                   flags, C.c_void_p(out.data_ptr()),
                                   C.c_void_p(torch.cuda.current_stream().cuda_stream))
    if rc != 0:
        return None
    torch.cuda.synchronize()
    return out


def lay(a, b):
    return ((a + 1) << 12) | ((b + 1) << 16)


def build_cases():
    """List of (name, thunk -> (out, ref)); every case is independent so that a faulting one can be skipped."""
    from focal_b200 import _cabi
    lib = _cabi.load_bringup()
    cases = []
    K = 32

    def mk(seed, *shape):
        g = torch.Generator(device="cuda").manual_seed(seed)
        return tf32_round(torch.randn(*shape, device="cuda", generator=g))

    def add(name, fn):
        cases.append((name, fn))

    for N in (64, 128):
        def u1(N=N, a_atom=16, b_img=None, a_lay=2, b_lay=2, a_par=(16, 1024, 32), b_par=(16, 1024, 32), a_inter=False):
            A, Bm = mk(1, 128, K), mk(2, N, K)
            ai = img_interleaved(A) if a_inter else img_rows(A, a_atom)
            bi = b_img(Bm)
            return run(lib, ai, bi, idesc(128, N), *a_par, *b_par, 4, N, TF32 | lay(a_lay, b_lay)), A.double() @ Bm.double().T
        add(f"U1 N={N} B K-major layout2 (16B atoms) sbo1024", lambda u1=u1: u1(b_img=lambda m: img_rows(m, 16)))
        for sbo in (512, 1024):
            add(f"U1 N={N} B K-major layout1 (32B atoms) sbo{sbo}",
                lambda u1=u1, sbo=sbo: u1(b_img=lambda m: img_rows(m, 32), b_lay=1, b_par=(16, sbo, 32)))
            add(f"U1 N={N} A+B K-major layout1 (32B atoms) sbo{sbo}",
                lambda u1=u1, sbo=sbo: u1(a_atom=32, a_lay=1, a_par=(16, sbo, 32), b_img=lambda m: img_rows(m, 32), b_lay=1,
                                          b_par=(16, sbo, 32)))
        g8 = K // 4 * 128
        add(f"U1 N={N} B K-major layout0 interleaved lbo128 sbo{g8}",
            lambda u1=u1, g8=g8: u1(b_img=img_interleaved, b_lay=0, b_par=(128, g8, 256)))
        add(f"U1 N={N} B K-major layout0 interleaved lbo{g8} sbo128",
            lambda u1=u1, g8=g8: u1(b_img=img_interleaved, b_lay=0, b_par=(g8, 128, 256)))
        add(f"U1 N={N} A+B K-major layout0 interleaved",
            lambda u1=u1, g8=g8: u1(a_inter=True, a_lay=0, a_par=(128, g8, 256), b_img=img_interleaved, b_lay=0,
                                    b_par=(128, g8, 256)))

    Kj = 32
    for Nd in (64, 256):
        for a_mode, a_tag in ((0, "A smem"), (2, "A tmem")):
            def u2(Nd=Nd, a_mode=a_mode, z_img=None, b_lay=2, b_par=None):
                W, Z = mk(3, 128, Kj), mk(4, Kj, Nd)
                ai = img_rows(W, 16) if a_mode == 0 else W.contiguous().view(torch.uint8).reshape(-1)
                ap = (16, 1024, 32) if a_mode == 0 else (0, 0, 0)
                out = run(lib, ai, z_img(Z), idesc(128, Nd, 0, 1), *ap, *b_par, Kj // 8, Nd, a_mode | TF32 | lay(2, b_lay))
                return out, W.double() @ Z.double()
            add(f"U2 Nd={Nd} {a_tag} Z MN layout2 (16B atoms) lbo=Kj*128 sbo1024 kstep1024",
                lambda u2=u2: u2(z_img=lambda m: img_rows(m, 16), b_lay=2, b_par=(Kj * 128, 1024, 1024)))
            for sbo in (512, 1024):
                add(f"U2 Nd={Nd} {a_tag} Z MN layout1 (32B atoms) lbo=Kj*128 sbo{sbo} kstep1024",
                    lambda u2=u2, sbo=sbo: u2(z_img=lambda m: img_rows(m, 32), b_lay=1, b_par=(Kj * 128, sbo, 1024)))
            add(f"U2 Nd={Nd} {a_tag} Z MN layout1 (32B atoms) lbo512 sbo=Kj*128 (swapped)",
                lambda u2=u2: u2(z_img=lambda m: img_rows(m, 32), b_lay=1, b_par=(512, Kj * 128, 1024)))
            grp = Nd // 4 * 128
            add(f"U2 Nd={Nd} {a_tag} Z MN layout0 interleaved sbo128 lbo{grp}",
                lambda u2=u2, grp=grp: u2(z_img=img_interleaved, b_lay=0, b_par=(grp, 128, grp)))
            add(f"U2 Nd={Nd} {a_tag} Z MN layout0 interleaved lbo128 sbo{grp}",
                lambda u2=u2, grp=grp: u2(z_img=img_interleaved, b_lay=0, b_par=(128, grp, grp)))
            add(f"U2 Nd={Nd} {a_tag} Z MN layout0 plain rows lbo=Kj*128 sbo1024",
                lambda u2=u2: u2(z_img=lambda m: img_rows(m, 0), b_lay=0, b_par=(Kj * 128, 1024, 1024)))
    return cases


def child(start):
    cases = build_cases()
    for i in range(start, len(cases)):
        name, fn = cases[i]
        print(f"BEGIN {i}", flush=True)
        out, ref = fn()
        if out is None:
            print(f"CASE {i} {name:78s} launch refused", flush=True)
            continue
        err = float((out.double() - ref).abs().max())
        nz = float((out != 0).float().mean())
        print(f"CASE {i} {name:78s} max|err| = {err:9.3e} nonzero = {nz:.2f} {'OK' if err < 1e-3 else '--'}", flush=True)
    print("DONE", flush=True)


def main():
    import subprocess
    if len(sys.argv) > 2 and sys.argv[1] == "--child":
        return child(int(sys.argv[2]))
    start, guard = 0, 0
    while guard < 80:
        guard += 1
        res = subprocess.run([sys.executable, os.path.abspath(__file__), "--child", str(start)], capture_output=True, text=True)
        last_begin, done = None, False
        for line in res.stdout.splitlines():
            if line.startswith("CASE "):
                print(line[5:])
            elif line.startswith("BEGIN "):
                last_begin = int(line.split()[1])
            elif line == "DONE":
                done = True
        if done:
            break
        if last_begin is None:
            print("child failed before the first case:", res.stderr[-400:])
            break
        err = [l for l in res.stderr.splitlines() if "CUDA error" in l or "rror" in l][-1:] or ["?"]
        print(f"{last_begin} FAULT ({err[0].strip()[:90]})")
        start = last_begin + 1


if __name__ == "__main__":
    main()
