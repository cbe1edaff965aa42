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


def report(name, out, ref):
    if out is None:
        print(f"{name:75s} launch refused")
        return
    err = float((out.double() - ref).abs().max())
    nz = float((out != 0).float().mean())
    print(f"{name:75s} max|err| = {err:9.3e}   nonzero = {nz:.2f}   {'OK' if err < 1e-3 else '--'}")


def main():
    from focal_b200 import _cabi
    lib = _cabi.load_bringup()
    g = torch.Generator(device="cuda").manual_seed(7)
    K = 32
    A = tf32_round(torch.randn(128, K, device="cuda", generator=g))
    a16 = img_rows(A, 16)
    print("== UMMA #1: S = A B^T, both K-major (A: SWIZZLE_128B from smem unless stated)")
    for N in (64, 128):
        Bm = tf32_round(torch.randn(N, K, device="cuda", generator=g))
        ref = A.double() @ Bm.double().T
        report(f"N={N} B K-major layout2 (16B atoms) sbo1024", run(lib, a16, img_rows(Bm, 16), idesc(128, N), 16, 1024, 32, 16, 1024, 32, 4, N, TF32 | lay(2, 2)), ref)
        for sbo in (512, 1024):
            report(f"N={N} B K-major layout1 (32B atoms) sbo{sbo}", run(lib, a16, img_rows(Bm, 32), idesc(128, N), 16, 1024, 32, 16, sbo, 32, 4, N, TF32 | lay(2, 1)), ref)
            report(f"N={N} A+B K-major layout1 (32B atoms) sbo{sbo}", run(lib, img_rows(A, 32), img_rows(Bm, 32), idesc(128, N), 16, sbo, 32, 16, sbo, 32, 4, N, TF32 | lay(1, 1)), ref)
        # no swizzle: core matrices [row/8][k/4][8][16 B]: LBO = next K unit (128 B), SBO = next row group (K/4 * 128 B)
        report(f"N={N} B K-major layout0 interleaved lbo128 sbo{K // 4 * 128}", run(lib, a16, img_interleaved(Bm), idesc(128, N), 16, 1024, 32, 128, K // 4 * 128, 256, 4, N, TF32 | lay(2, 0)), ref)
        report(f"N={N} B K-major layout0 interleaved (lbo/sbo swapped)", run(lib, a16, img_interleaved(Bm), idesc(128, N), 16, 1024, 32, K // 4 * 128, 128, 256, 4, N, TF32 | lay(2, 0)), ref)
        report(f"N={N} A+B K-major layout0 interleaved", run(lib, img_interleaved(A), img_interleaved(Bm), idesc(128, N), 128, K // 4 * 128, 256, 128, K // 4 * 128, 256, 4, N, TF32 | lay(0, 0)), ref)

    print("== UMMA #2: O = W Z, W K-major (smem, SWIZZLE_128B) or in TMEM, Z [Kj, Nd] MN-major")
    Kj = 32
    W = tf32_round(torch.randn(128, Kj, device="cuda", generator=g))
    w16 = img_rows(W, 16)
    w_raw = W.contiguous().view(torch.uint8).reshape(-1)
    for Nd in (64, 256):
        Z = tf32_round(torch.randn(Kj, Nd, device="cuda", generator=g))
        ref = W.double() @ Z.double()
        for a_mode, a_img, a_tag in ((0, w16, "A smem"), (2, w_raw, "A tmem")):
            ap = (16, 1024, 32) if a_mode == 0 else (0, 0, 0)
            idn = idesc(128, Nd, 0, 1)
            report(f"Nd={Nd} {a_tag} Z MN layout2 (16B atoms) lbo=Kj*128 sbo1024 kstep1024",
                   run(lib, a_img, img_rows(Z, 16), idn, *ap, Kj * 128, 1024, 1024, Kj // 8, Nd, a_mode | TF32 | lay(2, 2)), ref)
            for sbo in (512, 1024):
                report(f"Nd={Nd} {a_tag} Z MN layout1 (32B atoms) lbo=Kj*128 sbo{sbo} kstep1024",
                       run(lib, a_img, img_rows(Z, 32), idn, *ap, Kj * 128, sbo, 1024, Kj // 8, Nd, a_mode | TF32 | lay(2, 1)), ref)
            report(f"Nd={Nd} {a_tag} Z MN layout1 (32B atoms) lbo/sbo swapped",
                   run(lib, a_img, img_rows(Z, 32), idn, *ap, 512, Kj * 128, 1024, Kj // 8, Nd, a_mode | TF32 | lay(2, 1)), ref)
            # no swizzle: [j/8][n/4][8][16 B]: MN units 128 B apart, 8-row K groups (Nd/4)*128 B apart
            grp = Nd // 4 * 128
            report(f"Nd={Nd} {a_tag} Z MN layout0 interleaved sbo128 lbo{grp}",
                   run(lib, a_img, img_interleaved(Z), idn, *ap, grp, 128, grp, Kj // 8, Nd, a_mode | TF32 | lay(2, 0)), ref)
            report(f"Nd={Nd} {a_tag} Z MN layout0 interleaved lbo128 sbo{grp}",
                   run(lib, a_img, img_interleaved(Z), idn, *ap, 128, grp, grp, Kj // 8, Nd, a_mode | TF32 | lay(2, 0)), ref)
            report(f"Nd={Nd} {a_tag} Z MN layout0 plain rows [Nd/32][Kj][128B] lbo=Kj*128 sbo1024",
                   run(lib, a_img, img_rows(Z, 0), idn, *ap, Kj * 128, 1024, 1024, Kj // 8, Nd, a_mode | TF32 | lay(2, 0)), ref)

    print("== bf16 control: MN-major second GEMM with the shared SWIZZLE_128B tile (must be OK)")
    Wb = torch.randn(128, 64, device="cuda", generator=g).to(torch.bfloat16)
    Zb = torch.randn(64, 128, device="cuda", generator=g).to(torch.bfloat16)

    def img16(mat):
        R, Kc = mat.shape
        kb = Kc // 64
        x = mat.reshape(R, kb, 8, 8).permute(1, 0, 2, 3).contiguous()
        r = torch.arange(R, device=mat.device)
        c = torch.arange(8, device=mat.device)
        src = c[None, :] ^ (r[:, None] & 7)
        return torch.gather(x, 2, src[None, :, :, None].expand(kb, R, 8, 8)).contiguous().view(torch.uint8).reshape(-1)
    out = run(lib, img16(Wb), img16(Zb), idesc(128, 128, 0, 1, fmt=1), 16, 1024, 32, 64 * 128, 1024, 2048, 4, 128, 1)
    report("bf16 W smem, Z MN layout2", out, Wb.double() @ Zb.double())


if __name__ == "__main__":
    main()
