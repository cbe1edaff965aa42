"""GPU: pin the tcgen05 descriptor encodings the Gram kernels rely on (through the C ABI probe)."""
import ctypes as C

import pytest
import torch

pytestmark = pytest.mark.gpu


def _require_cuda():
    assert torch.cuda.is_available(), "GPU tests selected (-m gpu) but no CUDA device is visible"


def swizzled_image(mat: torch.Tensor) -> torch.Tensor:
    """[R, K] bf16 (K multiple of 64, R multiple of 8) -> K-block-major SWIZZLE_128B byte image (what the prologue
    writes to HBM and TMA copies verbatim): [K/64][R][128 B], 16-byte chunk c of row r stored at chunk c ^ (r & 7)."""
    R, K = mat.shape
    kb = K // 64
    x = mat.reshape(R, kb, 8, 8).permute(1, 0, 2, 3).contiguous()          # [kb, R, chunk, 8]
    r = torch.arange(R, device=mat.device)
    c = torch.arange(8, device=mat.device)
    src = (c[None, :] ^ (r[:, None] & 7))                                  # out[.., r, c'] = in[.., r, c' ^ (r&7)]
    out = torch.gather(x, 2, src[None, :, :, None].expand(kb, R, 8, 8))
    return out.contiguous().view(torch.uint8).reshape(-1)


def idesc(M, N, a_mn=0, b_mn=0, fmt=1):
    return (1 << 4) | (fmt << 7) | (fmt << 10) | (a_mn << 15) | (b_mn << 16) | ((N >> 3) << 17) | ((M >> 4) << 24)


def run_probe(a_img, b_img, idsc, a_lbo, a_sbo, a_kstep, b_lbo, b_sbo, b_kstep, ksteps, ncols, a_via_st=0):
    from focal_b200 import _cabi
    lib = _cabi.load_bringup()
    out = torch.full((128, ncols), float("nan"), device="cuda", dtype=torch.float32)
    rc = lib.focal_b200_debug_umma(C.c_void_p(a_img.data_ptr()), a_img.numel(), C.c_void_p(b_img.data_ptr()),
                                   b_img.numel(), idsc, a_lbo, a_sbo, a_kstep, b_lbo, b_sbo, b_kstep, ksteps, ncols,
                                   a_via_st, C.c_void_p(out.data_ptr()),
                                   C.c_void_p(torch.cuda.current_stream().cuda_stream))
    assert rc == 0, _cabi.strerror(rc)
    torch.cuda.synchronize()
    return out


@pytest.mark.parametrize("N", [64, 128])
@pytest.mark.parametrize("a_via_st", [0, 1])
def test_kmajor_gram_tile(N, a_via_st):
    """UMMA #1: S = A B^T with both operands K-major, SWIZZLE_128B, SBO = 1024, K step = 32 bytes."""
    _require_cuda()
    g = torch.Generator(device="cuda").manual_seed(1)
    A = torch.randn(128, 64, device="cuda", generator=g).to(torch.bfloat16)
    Bm = torch.randn(N, 64, device="cuda", generator=g).to(torch.bfloat16)
    out = run_probe(swizzled_image(A), swizzled_image(Bm), idesc(128, N), 16, 1024, 32, 16, 1024, 32, 4, N, a_via_st)
    ref = A.float() @ Bm.float().T
    assert torch.allclose(out, ref, rtol=1e-5, atol=1e-4), float((out - ref).abs().max())


@pytest.mark.parametrize("Nd", [64, 128, 256])
def test_mnmajor_second_gemm(Nd):
    """UMMA #2: O = W Z with W K-major (written by threads, st.shared) and Z MN-major: the SAME swizzled tile that
    served as the K-major B operand of UMMA #1.  LBO = bytes between 64-column groups, SBO = 1024, K step = 2048."""
    _require_cuda()
    g = torch.Generator(device="cuda").manual_seed(2)
    Kj = 64
    W = torch.randn(128, Kj, device="cuda", generator=g).to(torch.bfloat16)
    Z = torch.randn(Kj, Nd, device="cuda", generator=g).to(torch.bfloat16)
    out = run_probe(swizzled_image(W), swizzled_image(Z), idesc(128, Nd, 0, 1), 16, 1024, 32, Kj * 128, 1024, 2048,
                    Kj // 16, Nd, 1)
    ref = W.float() @ Z.float()
    assert torch.allclose(out, ref, rtol=1e-5, atol=1e-4), float((out - ref).abs().max())


@pytest.mark.parametrize("N,b_mn", [(64, 0), (128, 0), (128, 1), (256, 1)])
def test_a_operand_from_tensor_memory(N, b_mn):
    """TS mode: A lives in TMEM (lane = row, 32-bit column j = bf16 elements 2j, 2j+1), written with tcgen05.st;
    B is the usual swizzled smem tile, K-major (UMMA #1) or MN-major (UMMA #2)."""
    _require_cuda()
    g = torch.Generator(device="cuda").manual_seed(3)
    K = 64
    A = torch.randn(128, K, device="cuda", generator=g).to(torch.bfloat16)
    a_img = A.contiguous().view(torch.uint8).reshape(-1)
    if b_mn:
        Z = torch.randn(K, N, device="cuda", generator=g).to(torch.bfloat16)
        out = run_probe(a_img, swizzled_image(Z), idesc(128, N, 0, 1), 0, 0, 0, K * 128, 1024, 2048, K // 16, N, 2)
        ref = A.float() @ Z.float()
    else:
        Bm = torch.randn(N, K, device="cuda", generator=g).to(torch.bfloat16)
        out = run_probe(a_img, swizzled_image(Bm), idesc(128, N), 0, 0, 0, 16, 1024, 32, K // 16, N, 2)
        ref = A.float() @ Bm.float().T
    assert torch.allclose(out, ref, rtol=1e-5, atol=1e-4), float((out - ref).abs().max())


def swizzled_image32(mat: torch.Tensor) -> torch.Tensor:
    """Same byte layout for 32-bit elements: [R, K] fp32 (K multiple of 32) -> [K/32][R][128 B], chunks of 4 elements."""
    R, K = mat.shape
    kb = K // 32
    x = mat.reshape(R, kb, 8, 4).permute(1, 0, 2, 3).contiguous()
    r = torch.arange(R, device=mat.device)
    c = torch.arange(8, device=mat.device)
    src = (c[None, :] ^ (r[:, None] & 7))
    out = torch.gather(x, 2, src[None, :, :, None].expand(kb, R, 8, 4))
    return out.contiguous().view(torch.uint8).reshape(-1)


def tf32_round(x: torch.Tensor) -> torch.Tensor:
    """Round-to-nearest to 10 mantissa bits (what cvt.rna.tf32.f32 does), so the tensor core's truncation is exact."""
    i = x.contiguous().view(torch.int32)
    return ((i + 0x1000) & ~0x1FFF).view(torch.float32)


TF32 = 0x100


@pytest.mark.parametrize("N", [64, 96, 128])
def test_tf32_kmajor_gram_tile(N):
    """kind::tf32, both operands K-major: 32 elements per 128-byte swizzle row, K step = 8 elements = 32 bytes."""
    _require_cuda()
    g = torch.Generator(device="cuda").manual_seed(4)
    A = tf32_round(torch.randn(128, 64, device="cuda", generator=g))
    Bm = tf32_round(torch.randn(N, 64, device="cuda", generator=g))
    # two 32-element K blocks: block stride = rows * 128 bytes; 8 K steps of 8 elements, the probe advances the
    # descriptors linearly, so give it one block (4 steps) per call and add the partial products
    ref = A.double() @ Bm.double().T
    out = torch.zeros(128, N, device="cuda")
    for kb in range(2):
        a_img = swizzled_image32(A[:, kb * 32:(kb + 1) * 32].contiguous())
        b_img = swizzled_image32(Bm[:, kb * 32:(kb + 1) * 32].contiguous())
        out += run_probe(a_img, b_img, idesc(128, N, fmt=2), 16, 1024, 32, 16, 1024, 32, 4, N if N % 32 == 0 else 128,
                         TF32)[:, :N]
    assert torch.allclose(out.double(), ref, rtol=1e-5, atol=1e-4), float((out.double() - ref).abs().max())


def test_tf32_a_from_tensor_memory_kmajor_b():
    """kind::tf32 with A in tensor memory (one 32-bit column per element, 8 columns per K step) and a K-major B tile.
    (Measured while bringing this up: an MN-major B descriptor with kind::tf32 returns all zeros for every LBO / K-step
    combination tried, so a TF32 mode cannot re-use the K-major tile as the second GEMM's operand the way the bf16
    kernels do -- it needs a transposed operand copy.  See DESIGN.md, "Not built yet".)"""
    _require_cuda()
    g = torch.Generator(device="cuda").manual_seed(6)
    A = tf32_round(torch.randn(128, 32, device="cuda", generator=g))
    Bm = tf32_round(torch.randn(64, 32, device="cuda", generator=g))
    out = run_probe(A.contiguous().view(torch.uint8).reshape(-1), swizzled_image32(Bm), idesc(128, 64, fmt=2), 0, 0, 0,
                    16, 1024, 32, 4, 64, 2 | TF32)
    ref = A.double() @ Bm.double().T
    assert torch.allclose(out.double(), ref, rtol=1e-5, atol=1e-4), float((out.double() - ref).abs().max())
