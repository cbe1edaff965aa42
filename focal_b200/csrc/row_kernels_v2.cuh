// Row kernels, second generation: ONE WARP PER (row, tensor).
//
// prologue_v2_kernel / finalize_v2_kernel cover the same shapes as the vectorised kernels of row_kernels_fast.cuh
// (D even, D/2 a multiple of 32 up to 128, standard topology, S in {1, 2, 4}) and compute the same quantities, but a block
// holds 4 consecutive rows x all 2M tensors as 4 * 2M warps, each warp owning one tensor of one row:
//   * four times as many, four times shorter dependent chains per row: the kernels were latency-bound (ncu: issue slots
//     53 % / 74 % busy with 32 warps per SM, 17 clk per instruction per warp), not bandwidth-bound;
//   * nothing loops over the plan's tensor / operand tables inside a warp, so the per-tensor addresses are computed once;
//   * the reductions whose inputs are available together are batched (warp_sum_n: interleaved butterflies);
//   * the intra-sequence distances of the prologue are computed once per unordered pair (6 instead of 12 per sequence);
//   * the prologue stores the squared norms it needs anyway, finalize reads them instead of re-reducing.
// Per (row, tensor) the warp touches: 2 x 16-byte loads of the row per lane, the operand stores, and (finalize) its rows of
// the two accumulator sets -- every global access is a coalesced 8/16-byte vector.
#pragma once
#include "peer.cuh"
#include "plan.h"
#include "ptx.cuh"
#include "row_kernels.cuh"
#include "row_kernels_fast.cuh"

namespace fb {

// all-reduce of N independent values over the warp; the N butterflies are interleaved (latency of one)
template <int N>
__device__ __forceinline__ void warp_sum_n(float (&v)[N]) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
    for (int k = 0; k < N; ++k) v[k] += __shfl_xor_sync(0xffffffffu, v[k], o);
  }
}

// shared memory of both kernels: per-(row, tensor) scalars first (fixed offsets), then the staged rows
__host__ __device__ inline size_t row_v2_smem_bytes(int nT, int D) {
  return ((size_t)4 * nT * D + (size_t)4 * nT * 4) * sizeof(float);
}

// ---------------------------------------------------------------------------------------------------------
// prologue: norms, InfoNCE + temporal operands, orthogonality terms, intra-sequence mean distances m_II
// ---------------------------------------------------------------------------------------------------------
template <int VW, int PREC>
__global__ void __launch_bounds__(1024, 1) prologue_v2_kernel(const __grid_constant__ Plan p,
                                                              const __grid_constant__ FeatPtrs f,
                                                              const __grid_constant__ PeerWs pw,
                                                              uint8_t* __restrict__ ws, int fuse_intra) {
  extern __shared__ float smem_f[];
  const int nT = p.nT, D = p.D, d = p.d, S = p.S;
  // warp index through a shuffle: provably warp-uniform, so that every branch on (row, tensor) state below is uniform for
  // the compiler too and the shuffle reductions inside them need no WARPSYNC / collective wrappers
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  const int r = warp / nT, t = warp - r * nT;                 // row within the block, tensor
  const int row_lo = p.local_rows ? p.seq0 * S : 0, row_hi = p.local_rows ? p.seq1 * S : p.B;
  const int i = row_lo + blockIdx.x * 4 + r;
  const bool live = i < row_hi;
  float* nrm2 = smem_f;                                       // [4][nT][2] squared norms (shared half, private half)
  float* part = nrm2 + 4 * nT * 2;                            // [4][nT] intra-sequence partial sums
  float* red = part + 4 * nT;                                 // [4][nT] orthogonality partial sums
  float* xs = red + 4 * nT;                                   // [4][nT][D] raw rows (16 * nT floats in: 16-byte aligned)
  const int c0 = VW * lane;
  const int I = i / S, s = i - I * S;
  const bool owned = live && I >= p.seq0 && I < p.seq1;
  const bool tmp_on = (p.terms & FOCAL_TERM_TEMPORAL) != 0;
  float sh[VW], pr[VW], rsh[VW], rpr[VW];
  float acc_orth = 0.f, na = 1.f, nb = 1.f;
  if (live) {
    const float* src = feat_base(p, f, ws, t) + feat_row_off(p, i);
    ld_frag<VW>(src + c0, sh);
    ld_frag<VW>(src + d + c0, pr);
    float* mine = xs + ((size_t)r * nT + t) * D;
    st_frag<VW>(mine + c0, sh);
    st_frag<VW>(mine + d + c0, pr);
    float q4[4] = {0.f, 0.f, 0.f, 0.f};                       // |shared|^2, |private|^2, |rounded row|^2, shared . private
#pragma unroll
    for (int e = 0; e < VW; ++e) {
      q4[0] = fmaf(sh[e], sh[e], q4[0]);
      q4[1] = fmaf(pr[e], pr[e], q4[1]);
      q4[3] = fmaf(sh[e], pr[e], q4[3]);
      rsh[e] = op_round_t<PREC>(sh[e]);
      rpr[e] = op_round_t<PREC>(pr[e]);
      q4[2] += tile_sq(PREC, sh[e]) + tile_sq(PREC, pr[e]);
    }
    warp_sum_n<4>(q4);
    na = q4[0]; nb = q4[1];
    if (lane == 0) {
      nrm2[(r * nT + t) * 2] = na;
      nrm2[(r * nT + t) * 2 + 1] = nb;
      *reinterpret_cast<float2*>(ws + p.nrm_off + ((uint64_t)t * p.Bpad + i) * 8) = make_float2(na, nb);
    }
    // ---- InfoNCE operands: x / max(|x|, eps) * sqrt(log2 e / T), position-major rows
    if (p.terms & FOCAL_TERM_NCE) {
      const float fa = p.alpha * fminf(rsqrtf(na), 1.f / kNceEps), fb2 = p.alpha * fminf(rsqrtf(nb), 1.f / kNceEps);
      float zs[VW], zp[VW];
#pragma unroll
      for (int e = 0; e < VW; ++e) { zs[e] = sh[e] * fa; zp[e] = pr[e] * fb2; }
      const uint64_t rowN = (uint64_t)s * p.bpad + I, rowsNce = (uint64_t)S * p.bpad;
      const uint64_t off_s = p.ops[2 * t].off, off_p = p.ops[2 * t + 1].off;
      const int kh = (PREC == FOCAL_PREC_FP32) ? p.ops[2 * t].kb / 2 : p.ops[2 * t].kb;
      for (int rk = 0; rk < pw.world; ++rk) {
        uint8_t* w = pw.ws[rk];
        st_operand_p<VW, PREC>(w + off_s, rowsNce, rowN, c0, zs, kh);
        st_operand_p<VW, PREC>(w + off_p, rowsNce, rowN, c0, zp, kh);
        if (VW & 1) {                                         // d = 32 or 96: zero the unused half of the last K block
          const float z1[1] = {0.f};
          st_operand_p<1, PREC>(w + off_s, rowsNce, rowN, d + lane, z1, kh);
          st_operand_p<1, PREC>(w + off_p, rowsNce, rowN, d + lane, z1, kh);
        }
      }
    }
    // ---- temporal operands (raw rows, natural order) + squared norm of what the tiles will see
    if (tmp_on) {
      const int khf = (PREC == FOCAL_PREC_FP32) ? p.kbFull / 2 : p.kbFull;
      const uint64_t xoff = p.xt_off + (uint64_t)t * p.kbFull * p.Bpad * 128;
      for (int rk = 0; rk < pw.world; ++rk) {
        uint8_t* xt = pw.ws[rk] + xoff;
        st_operand_p<VW, PREC>(xt, (uint64_t)p.Bpad, (uint64_t)i, c0, sh, khf);
        st_operand_p<VW, PREC>(xt, (uint64_t)p.Bpad, (uint64_t)i, d + c0, pr, khf);
      }
      if (lane < pw.world) reinterpret_cast<float*>(pw.ws[lane] + p.sq_off)[(uint64_t)t * p.Bpad + i] = q4[2];
    }
    // ---- orthogonality (loss.py:96-104), pair (shared_t, private_t): lane-local dot product, already reduced
    if (owned && (p.terms & FOCAL_TERM_ORTH))
      acc_orth = fmaxf(q4[3] * rsqrtf((na + kOrthEps) * (nb + kOrthEps)), 0.f);
  }
  __syncthreads();
  const bool intra = fuse_intra && tmp_on && S > 1 && p.b > 1;
  if (live) {
    // ---- orthogonality, pairs (private_t, private_t') with t' > t of the same view
    if (owned && (p.terms & FOCAL_TERM_ORTH)) {
      const int vend = (t / p.M + 1) * p.M;
      for (int t2 = t + 1; t2 < vend; ++t2) {
        float v[VW];
        ld_frag<VW>(xs + ((size_t)r * nT + t2) * D + d + c0, v);
        float dot = 0.f;
#pragma unroll
        for (int e = 0; e < VW; ++e) dot = fmaf(pr[e], v[e], dot);
        dot = warp_sum(dot);
        acc_orth += fmaxf(dot * rsqrtf((nb + kOrthEps) * (nrm2[(r * nT + t2) * 2 + 1] + kOrthEps)), 0.f);
      }
    }
    // ---- intra-sequence distances of the rounded rows, each unordered pair once: row s takes (s, s+1 mod S) and,
    // for S = 4, rows 0 and 1 also take (s, s+2); S = 2: row 0 takes the only pair
    if (intra) {
      const int r0 = r - s;                                   // block row of position 0 of this sequence
      int j1 = -1, j2 = -1;
      if (S == 4) { j1 = (s + 1) & 3; if (s < 2) j2 = s + 2; }
      else if (s == 0) j1 = 1;
      float d2[2] = {0.f, 0.f};
      if (j1 >= 0) {
        const float* xo = xs + ((size_t)(r0 + j1) * nT + t) * D;
        float osh[VW], opr[VW];
        ld_frag<VW>(xo + c0, osh);
        ld_frag<VW>(xo + d + c0, opr);
#pragma unroll
        for (int e = 0; e < VW; ++e) {
          const float a = rsh[e] - op_round_t<PREC>(osh[e]), b = rpr[e] - op_round_t<PREC>(opr[e]);
          d2[0] = fmaf(a, a, fmaf(b, b, d2[0]));
        }
      }
      if (j2 >= 0) {
        const float* xo = xs + ((size_t)(r0 + j2) * nT + t) * D;
        float osh[VW], opr[VW];
        ld_frag<VW>(xo + c0, osh);
        ld_frag<VW>(xo + d + c0, opr);
#pragma unroll
        for (int e = 0; e < VW; ++e) {
          const float a = rsh[e] - op_round_t<PREC>(osh[e]), b = rpr[e] - op_round_t<PREC>(opr[e]);
          d2[1] = fmaf(a, a, fmaf(b, b, d2[1]));
        }
      }
      warp_sum_n<2>(d2);
      if (lane == 0) part[r * nT + t] = sqrtf(d2[0]) + sqrtf(d2[1]);
    }
  }
  if (lane == 0) red[warp] = acc_orth;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s2 = 0.f;
    for (int w = 0; w < 4 * nT; ++w) s2 += red[w];            // fixed order: deterministic
    float* p1 = reinterpret_cast<float*>(ws + p.part1_off) + (size_t)blockIdx.x * 4;
    p1[0] = s2 / (float)p.B; p1[1] = 0.f; p1[2] = 0.f;
  }
  if (intra && live && lane == 0) {
    const int r0 = r - s;
    float m = 0.f;
    for (int j = 0; j < S; ++j) m += part[(r0 + j) * nT + t];
    m = 2.f * m / (float)(S * S - S);
    for (int rk = 0; rk < pw.world; ++rk)
      reinterpret_cast<float*>(pw.ws[rk] + p.mintra_off)[(uint64_t)t * p.Bpad + i] = m;
  }
  if (pw.world > 1) peer_epoch_bump(p, ws);      // operands of the owned rows are out
}

// ---------------------------------------------------------------------------------------------------------
// finalize: gradient row of one tensor = temporal part + the InfoNCE operands of the tensor + its orthogonality pairs
// ---------------------------------------------------------------------------------------------------------
template <int VW, int PREC>
__global__ void __launch_bounds__(1024, 1) finalize_v2_kernel(const __grid_constant__ Plan p,
                                                              const __grid_constant__ FeatPtrs f,
                                                              const __grid_constant__ GradPtrs g,
                                                              const uint8_t* __restrict__ ws) {
  extern __shared__ float smem_f[];
  const int nT = p.nT, D = p.D, d = p.d, S = p.S;
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;    // warp-uniform
  const int r = warp / nT, t = warp - r * nT;
  const int i = p.seq0 * S + blockIdx.x * 4 + r;
  const bool live = i < p.seq1 * S;
  float* nrm2 = smem_f;                                       // [4][nT][2]
  float* xs = smem_f + 16 * nT;                               // [4][nT][D] (same layout as the prologue)
  const int c0 = VW * lane;
  float sh[VW], pr[VW];
  float na = 1.f, nb = 1.f;
  if (live) {
    const float* src = feat_base(p, f, ws, t) + feat_row_off(p, i);
    ld_frag<VW>(src + c0, sh);
    ld_frag<VW>(src + d + c0, pr);
    float* mine = xs + ((size_t)r * nT + t) * D;
    st_frag<VW>(mine + c0, sh);
    st_frag<VW>(mine + d + c0, pr);
    const float2 n2 = __ldg(reinterpret_cast<const float2*>(ws + p.nrm_off + ((uint64_t)t * p.Bpad + i) * 8));
    na = n2.x; nb = n2.y;
    if (lane == 0) { nrm2[(r * nT + t) * 2] = na; nrm2[(r * nT + t) * 2 + 1] = nb; }
  }
  __syncthreads();
  if (!live) return;
  const int I = i / S, s = i - I * S;
  float gsh[VW], gpr[VW];
#pragma unroll
  for (int e = 0; e < VW; ++e) { gsh[e] = 0.f; gpr[e] = 0.f; }

  // ---- temporal part: x~_i rho_i - (R X~)_i, then the exact intra-sequence pairs
  if ((p.terms & FOCAL_TERM_TEMPORAL) && p.b > 1 && S > 1) {
    const int Dp = p.kbFull * p.epb;
    const uint64_t ti = (uint64_t)t * p.Bpad + i;
    const int extra = __ldg(reinterpret_cast<const int32_t*>(ws + p.flag_tmp_off) + (uint64_t)t * (p.Bpad / kTileM) + i / kTileM);
    float rho = __ldg(reinterpret_cast<const float*>(ws + p.rho_off) + ti);
    int cnt = __ldg(reinterpret_cast<const int32_t*>(ws + p.cnt_off) + (uint64_t)t * p.bpad + I);
    const float* y = reinterpret_cast<const float*>(ws + p.dx_off) + ti * Dp;
    float ysh[VW], ypr[VW];
    ld_frag<VW>(y + c0, ysh);
    ld_frag<VW>(y + d + c0, ypr);
    for (int k = 1; k <= extra; ++k) {
      const float* y2 = reinterpret_cast<const float*>(ws + p.dx_off + k * p.dx2_delta) + ti * Dp;
      float zsh[VW], zpr[VW];
      ld_frag<VW>(y2 + c0, zsh);
      ld_frag<VW>(y2 + d + c0, zpr);
#pragma unroll
      for (int e = 0; e < VW; ++e) { ysh[e] += zsh[e]; ypr[e] += zpr[e]; }
      rho += __ldg(reinterpret_cast<const float*>(ws + p.rho_off + k * p.rho2_delta) + ti);
      cnt += __ldg(reinterpret_cast<const int32_t*>(ws + p.cnt_off + k * p.cnt2_delta) + (uint64_t)t * p.bpad + I);
    }
    float rsh[VW], rpr[VW];
#pragma unroll
    for (int e = 0; e < VW; ++e) {
      rsh[e] = op_round_t<PREC>(sh[e]); rpr[e] = op_round_t<PREC>(pr[e]);
      gsh[e] = p.w_rank * fmaf(rsh[e], rho, -ysh[e]);
      gpr[e] = p.w_rank * fmaf(rpr[e], rho, -ypr[e]);
    }
    // dL/dm_II = cnt / (b (b-1)), spread over S^2 - S ordered pairs, both orders; the shuffles run unconditionally
    // (cnt is uniform per sequence, but the compiler cannot know) and a zero coefficient switches the update off
    const float coef = p.w_rank * 2.f * (float)cnt / ((float)p.b * (float)(p.b - 1) * (float)(S * S - S));
    const int r0 = r - s;
    for (int jj = 1; jj < S; ++jj) {
      const int j = (s + jj) % S;
      const float* xo = xs + ((size_t)(r0 + j) * nT + t) * D;
      float osh[VW], opr[VW];
      ld_frag<VW>(xo + c0, osh);
      ld_frag<VW>(xo + d + c0, opr);
      float d2 = 0.f;
#pragma unroll
      for (int e = 0; e < VW; ++e) {
        osh[e] = rsh[e] - op_round_t<PREC>(osh[e]); opr[e] = rpr[e] - op_round_t<PREC>(opr[e]);
        d2 = fmaf(osh[e], osh[e], fmaf(opr[e], opr[e], d2));
      }
      d2 = warp_sum(d2);
      const float rr = d2 > 0.f ? coef * rsqrtf(d2) : 0.f;
#pragma unroll
      for (int e = 0; e < VW; ++e) { gsh[e] = fmaf(rr, osh[e], gsh[e]); gpr[e] = fmaf(rr, opr[e], gpr[e]); }
    }
  }

  // ---- InfoNCE: operands 2t (shared half) and 2t + 1 (private half) of this tensor
  if (p.terms & FOCAL_TERM_NCE) {
    const uint64_t rowN = (uint64_t)s * p.bpad + I;
    const float inv_tsn = 1.f / (p.T * (float)S * (float)(2 * p.b));
    const float inv_alpha = 1.f / p.alpha;
    float tmp[2][VW];
#pragma unroll
    for (int half = 0; half < 2; ++half) {
#pragma unroll
      for (int e = 0; e < VW; ++e) tmp[half][e] = 0.f;
      const OpDesc& op = p.ops[2 * t + half];
      const int wp = op.kb * p.epb;
      for (int u = 0; u < op.nuse; ++u) {
        const int q = op.use_prob[u], side = op.use_side[u];
        const ProbDesc& prb = p.probs[q];
        const OpDesc& po = p.ops[op.use_partner[u]];                  // partner operand: positive row p(k)
        const int ph = po.col0 ? 1 : 0;
        const float fp = p.alpha * fminf(rsqrtf(nrm2[(r * nT + po.tensor) * 2 + ph]), 1.f / kNceEps);
        float px[VW], acc[VW];
        ld_frag<VW>(xs + ((size_t)r * nT + po.tensor) * D + po.col0 + c0, px);
        const uint64_t arow = ((uint64_t)side * S * p.bpad + rowN) * wp + c0;
        ld_frag<VW>(reinterpret_cast<const float*>(ws + prb.dz_off) + arow, acc);
        const uint64_t fidx = ((uint64_t)q * S + s) * 2 + side;
        const int extra = __ldg(reinterpret_cast<const int32_t*>(ws + p.flag_nce_off) + fidx * (p.bpad / kTileM) + I / kTileM);
        for (int k = 1; k <= extra; ++k) {
          float acc2[VW];
          ld_frag<VW>(reinterpret_cast<const float*>(ws + prb.dz_off + k * p.dz2_delta) + arow, acc2);
#pragma unroll
          for (int e = 0; e < VW; ++e) acc[e] += acc2[e];
        }
        const float* rs = reinterpret_cast<const float*>(ws + p.rsum_off) + ((uint64_t)(q * S + s) * 2) * p.bpad;
        const float r_k = __ldg(rs + (uint64_t)side * p.bpad + I), r_p = __ldg(rs + (uint64_t)(1 - side) * p.bpad + I);
        // positive column in fp32 (masked out of the tiles): W_kp - 2 is a tiny difference when the positive dominates;
        // its logit is the one the row-sum tile of this row saw
        const float gpos = __ldg(reinterpret_cast<const float*>(ws + p.pos_off) + fidx * p.bpad + I);
        const float wkp = ex2_approx(gpos) * (__frcp_rn(r_k) + __frcp_rn(r_p));
        const float wq = prb.weight * inv_tsn * inv_alpha;
#pragma unroll
        for (int e = 0; e < VW; ++e)
          tmp[half][e] = fmaf(wq, fmaf(wkp - 2.f, op_round_t<PREC>(px[e] * fp), acc[e]), tmp[half][e]);
      }
    }
    float dots[2] = {0.f, 0.f};                         // d zh / d z = (I - zh zh^T) / n
#pragma unroll
    for (int e = 0; e < VW; ++e) { dots[0] = fmaf(tmp[0][e], sh[e], dots[0]); dots[1] = fmaf(tmp[1][e], pr[e], dots[1]); }
    warp_sum_n<2>(dots);
    const float ia = fminf(rsqrtf(na), 1.f / kNceEps), ib = fminf(rsqrtf(nb), 1.f / kNceEps);      // 1 / max(|z|, eps)
    const float da = dots[0] * ia * ia, db = dots[1] * ib * ib;
#pragma unroll
    for (int e = 0; e < VW; ++e) {
      gsh[e] = fmaf(fmaf(-da, sh[e], tmp[0][e]), ia, gsh[e]);
      gpr[e] = fmaf(fmaf(-db, pr[e], tmp[1][e]), ib, gpr[e]);
    }
  }

  // ---- orthogonality: (shared_t, private_t) and (private_t, private_t') for every other t' of the view
  if (p.terms & FOCAL_TERM_ORTH) {
    const float a = p.w_orth / (float)p.B;
    {
      float dot = 0.f;
#pragma unroll
      for (int e = 0; e < VW; ++e) dot = fmaf(sh[e], pr[e], dot);
      dot = warp_sum(dot);
      const float nu = na + kOrthEps, nv = nb + kOrthEps;
      const float inv_den = rsqrtf(nu * nv), cs = dot * inv_den;
      const float on = cs >= 0.f ? a : 0.f;                 // clamp_min passes gradient at equality
      const float ad = on * inv_den, au = -on * cs * __frcp_rn(nu), av = -on * cs * __frcp_rn(nv);
#pragma unroll
      for (int e = 0; e < VW; ++e) {
        gsh[e] = fmaf(ad, pr[e], fmaf(au, sh[e], gsh[e]));
        gpr[e] = fmaf(ad, sh[e], fmaf(av, pr[e], gpr[e]));
      }
    }
    const int v0 = (t / p.M) * p.M;
    for (int t2 = v0; t2 < v0 + p.M; ++t2) {
      if (t2 == t) continue;
      float v[VW];
      ld_frag<VW>(xs + ((size_t)r * nT + t2) * D + d + c0, v);
      float dot = 0.f;
#pragma unroll
      for (int e = 0; e < VW; ++e) dot = fmaf(pr[e], v[e], dot);
      dot = warp_sum(dot);
      const float nu = nb + kOrthEps, nv = nrm2[(r * nT + t2) * 2 + 1] + kOrthEps;
      const float inv_den = rsqrtf(nu * nv), cs = dot * inv_den;
      const float on = cs >= 0.f ? a : 0.f;
      const float ad = on * inv_den, au = -on * cs * __frcp_rn(nu);
#pragma unroll
      for (int e = 0; e < VW; ++e) gpr[e] = fmaf(ad, v[e], fmaf(au, pr[e], gpr[e]));
    }
  }
  float* out = grad_base(p, g, ws, t) + (size_t)i * D;
  st_frag<VW>(out + c0, gsh);
  st_frag<VW>(out + d + c0, gpr);
}

}  // namespace fb
