// Vectorised variants of the prologue / finalize row kernels for the common shapes:
//   D even, D/2 a multiple of 32 (D = 64, 128, 192, 256, and 512), standard topology (no "noPrivate"), S in {2, 4}.
// One warp per row, 4 consecutive rows (= whole sequences) per block.  Each lane owns VW = D/64 (1..4, or 8) consecutive columns
// of the shared half and the matching VW columns of the private half of every tensor, so all per-row dot products are
// lane-local multiplies + one shuffle reduction, every global access is a coalesced 8/16-byte vector, and the
// neighbouring windows of a sequence are read from shared memory instead of HBM.  The generic kernels in
// row_kernels.cuh remain the path for every other shape; both produce identical operand bytes.
#pragma once
#include "plan.h"
#include "ptx.cuh"
#include "row_kernels.cuh"

namespace fb {

template <int VW>
__device__ __forceinline__ void ld_frag(const float* p, float (&v)[VW]) {
  if (VW == 8) {
    const float4 t = *reinterpret_cast<const float4*>(p), u = *reinterpret_cast<const float4*>(p + 4);
    v[0] = t.x; v[1 % VW] = t.y; v[2 % VW] = t.z; v[3 % VW] = t.w;
    v[4 % VW] = u.x; v[5 % VW] = u.y; v[6 % VW] = u.z; v[7 % VW] = u.w;
  } else if (VW == 4) {
    const float4 t = *reinterpret_cast<const float4*>(p);
    v[0] = t.x; v[1] = t.y; v[2 % VW] = t.z; v[3 % VW] = t.w;
  } else if (VW == 2) {
    const float2 t = *reinterpret_cast<const float2*>(p);
    v[0] = t.x; v[1 % VW] = t.y;
  } else {
#pragma unroll
    for (int e = 0; e < VW; ++e) v[e] = p[e];
  }
}
template <int VW>
__device__ __forceinline__ void st_frag(float* p, const float (&v)[VW]) {
  if (VW == 8) {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1 % VW], v[2 % VW], v[3 % VW]);
    *reinterpret_cast<float4*>(p + 4) = make_float4(v[4 % VW], v[5 % VW], v[6 % VW], v[7 % VW]);
  } else if (VW == 4) {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2 % VW], v[3 % VW]);
  } else if (VW == 2) {
    *reinterpret_cast<float2*>(p) = make_float2(v[0], v[1 % VW]);
  } else {
#pragma unroll
    for (int e = 0; e < VW; ++e) p[e] = v[e];
  }
}

// Store VW consecutive bf16 operand elements starting at element e0 (multiple of VW) of a swizzled 128-byte-row
// operand array: K block e0/64, 16-byte chunk ((e0 % 64) / 8) ^ (row & 7).
template <int VW>
__device__ __forceinline__ void st_operand(uint8_t* op_base, uint64_t kstride_rows, uint64_t row, int e0,
                                           const float (&v)[VW]) {
  uint8_t* dst = op_base + ((uint64_t)(e0 >> 6) * kstride_rows + row) * 128 +
                 ((((uint32_t)(e0 & 63) >> 3) ^ (uint32_t)(row & 7)) << 4) + (e0 & 7) * 2;
  if (VW == 8) {
    *reinterpret_cast<uint4*>(dst) = make_uint4(pack_bf16x2(v[0], v[1 % VW]), pack_bf16x2(v[2 % VW], v[3 % VW]),
                                                pack_bf16x2(v[4 % VW], v[5 % VW]), pack_bf16x2(v[6 % VW], v[7 % VW]));
  } else if (VW == 4) {
    *reinterpret_cast<uint2*>(dst) = make_uint2(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2 % VW], v[3 % VW]));
  } else if (VW == 2) {
    *reinterpret_cast<uint32_t*>(dst) = pack_bf16x2(v[0], v[1 % VW]);
  } else {
#pragma unroll
    for (int e = 0; e < VW; ++e) {
      const int el = e0 + e;
      uint8_t* d1 = op_base + ((uint64_t)(el >> 6) * kstride_rows + row) * 128 +
                    ((((uint32_t)(el & 63) >> 3) ^ (uint32_t)(row & 7)) << 4) + (el & 7) * 2;
      *reinterpret_cast<uint16_t*>(d1) = (uint16_t)(pack_bf16x2(v[e], 0.f) & 0xffffu);
    }
  }
}

// Precision-aware operand store: bf16 tiles store bf16(v); split tiles (fp32 mode) store hi = bf16(v) into the hi image
// and lo = bf16(v - hi) into the lo image, `kh` K blocks further.
template <int VW, int PREC>
__device__ __forceinline__ void st_operand_p(uint8_t* op_base, uint64_t kstride_rows, uint64_t row, int e0,
                                             const float (&v)[VW], int kh) {
  st_operand<VW>(op_base, kstride_rows, row, e0, v);
  if (PREC == FOCAL_PREC_FP32) {
    float lo[VW];
#pragma unroll
    for (int e = 0; e < VW; ++e) lo[e] = v[e] - bf16_round(v[e]);
    st_operand<VW>(op_base + (uint64_t)kh * kstride_rows * 128, kstride_rows, row, e0, lo);
  }
}

__device__ __forceinline__ void stage_rows(const Plan& p, const FeatPtrs& f, const uint8_t* ws, int i, float* xs,
                                           int lane) {
  const int nv = p.D >> 2;                       // D % 4 == 0 on this path
  for (int t = 0; t < p.nT; ++t) {
    const float4* src = reinterpret_cast<const float4*>(feat_base(p, f, ws, t) + feat_row_off(p, i));
    float4* dst = reinterpret_cast<float4*>(xs + t * p.D);
    for (int c = lane; c < nv; c += 32) dst[c] = __ldg(src + c);
  }
}

// ---------------------------------------------------------------------------------------------------------
// fast prologue (+ fused intra-sequence means when S divides 4)
// ---------------------------------------------------------------------------------------------------------
template <int VW, int PREC>
__global__ void __launch_bounds__(128) prologue_fast_kernel(const __grid_constant__ Plan p,
                                                            const __grid_constant__ FeatPtrs f,
                                                            const __grid_constant__ PeerWs pw,
                                                            uint8_t* __restrict__ ws, int fuse_intra) {
  // Row-sharded jobs (p.local_rows): only the owned rows are processed, and everything the Gram kernels of ANY rank
  // read (operands, squared norms, m_II) is stored into every rank's workspace (pw.ws[r], NVLink peer stores).
  extern __shared__ float smem_f[];
  __shared__ float red[4][4];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row_lo = p.local_rows ? p.seq0 * p.S : 0, row_hi = p.local_rows ? p.seq1 * p.S : p.B;
  const int i = row_lo + blockIdx.x * 4 + warp;
  const int D = p.D, d = p.d;
  float* xs = smem_f + (size_t)warp * p.nT * D;
  float acc_orth = 0.f, acc_ps = 0.f, acc_pp = 0.f;
  const bool live = i < row_hi;
  if (live) stage_rows(p, f, ws, i, xs, lane);
  __syncthreads();
  if (live) {
    const int I = i / p.S, s = i % p.S;
    const uint64_t rowN = (uint64_t)s * p.bpad + I;
    const uint64_t rowsNce = (uint64_t)p.S * p.bpad;
    const int c0 = VW * lane;
    float* nrm = smem_f + (size_t)4 * p.nT * D + warp * 2 * kMaxT;     // [t][2] scale factors alpha / max(|.|, eps)

    for (int t = 0; t < p.nT; ++t) {
      float sh[VW], pr[VW];
      ld_frag<VW>(xs + t * D + c0, sh);
      ld_frag<VW>(xs + t * D + d + c0, pr);
      float a = 0.f, b = 0.f;
#pragma unroll
      for (int e = 0; e < VW; ++e) { a = fmaf(sh[e], sh[e], a); b = fmaf(pr[e], pr[e], b); }
      a = warp_sum(a); b = warp_sum(b);
      if (lane == 0) { nrm[2 * t] = a; nrm[2 * t + 1] = b; }
      if (p.terms & FOCAL_TERM_NCE) {
        const float fa = p.alpha * fminf(rsqrtf(a), 1.f / kNceEps), fb2 = p.alpha * fminf(rsqrtf(b), 1.f / kNceEps);
        float zs[VW], zp[VW];
#pragma unroll
        for (int e = 0; e < VW; ++e) { zs[e] = sh[e] * fa; zp[e] = pr[e] * fb2; }
        for (int r = 0; r < pw.world; ++r) {
          uint8_t* w = pw.ws[r];
          const int kh = (PREC == FOCAL_PREC_FP32) ? p.ops[2 * t].kb / 2 : p.ops[2 * t].kb;
          st_operand_p<VW, PREC>(w + p.ops[2 * t].off, rowsNce, rowN, c0, zs, kh);
          st_operand_p<VW, PREC>(w + p.ops[2 * t + 1].off, rowsNce, rowN, c0, zp, kh);
          if (VW & 1) {
            // d = 32 or 96: the last K block is half full -- its 32 padding columns must read as zeros
            const float z1[1] = {0.f};
            st_operand_p<1, PREC>(w + p.ops[2 * t].off, rowsNce, rowN, d + lane, z1, kh);
            st_operand_p<1, PREC>(w + p.ops[2 * t + 1].off, rowsNce, rowN, d + lane, z1, kh);
          }
        }
      }
      if (p.terms & FOCAL_TERM_TEMPORAL) {
        for (int r = 0; r < pw.world; ++r) {
          uint8_t* xt = pw.ws[r] + p.xt_off + (uint64_t)t * p.kbFull * p.Bpad * 128;
          const int khf = (PREC == FOCAL_PREC_FP32) ? p.kbFull / 2 : p.kbFull;
          st_operand_p<VW, PREC>(xt, (uint64_t)p.Bpad, (uint64_t)i, c0, sh, khf);
          st_operand_p<VW, PREC>(xt, (uint64_t)p.Bpad, (uint64_t)i, d + c0, pr, khf);
        }
        float sq = 0.f;
#pragma unroll
        for (int e = 0; e < VW; ++e) {
          sq += tile_sq(PREC, sh[e]) + tile_sq(PREC, pr[e]);
        }
        sq = warp_sum(sq);
        if (lane < pw.world) reinterpret_cast<float*>(pw.ws[lane] + p.sq_off)[(uint64_t)t * p.Bpad + i] = sq;
      }
    }
    __syncwarp();
    if (I >= p.seq0 && I < p.seq1) {
      if (p.terms & FOCAL_TERM_ORTH) {
        for (int k = 0; k < p.nOrth; ++k) {
          const OrthDesc& od = p.orth[k];
          float u[VW], v[VW];
          ld_frag<VW>(xs + od.tu * D + od.cu + c0, u);
          ld_frag<VW>(xs + od.tv * D + od.cv + c0, v);
          float dot = 0.f;
#pragma unroll
          for (int e = 0; e < VW; ++e) dot = fmaf(u[e], v[e], dot);
          dot = warp_sum(dot);
          const float nu = nrm[2 * od.tu + (od.cu ? 1 : 0)] + kOrthEps;
          const float nv = nrm[2 * od.tv + (od.cv ? 1 : 0)] + kOrthEps;
          acc_orth += fmaxf(dot * rsqrtf(nu * nv), 0.f);
        }
      }
    }
    // ---- intra-sequence mean distance of the rounded rows (the other rows of the sequence sit in this block)
    if (fuse_intra && (p.terms & FOCAL_TERM_TEMPORAL) && p.S > 1 && p.b > 1) {
      const int S = p.S;
      const int w0 = warp - s;                        // warp holding position 0 of this sequence
      // per-row sums sum_j delta(i, j); the S rows of the sequence are added up through shared memory below
      float* rowsum = smem_f + (size_t)4 * p.nT * D + 4 * 2 * kMaxT + warp * kMaxT;
      for (int t = 0; t < p.nT; ++t) {
        float sh[VW], pr[VW];
        ld_frag<VW>(xs + t * D + c0, sh);
        ld_frag<VW>(xs + t * D + d + c0, pr);
        float sum = 0.f;
        for (int j = 0; j < S; ++j) {
          if (j == s) continue;
          const float* xo = smem_f + (size_t)(w0 + j) * p.nT * D + t * D;
          float osh[VW], opr[VW];
          ld_frag<VW>(xo + c0, osh);
          ld_frag<VW>(xo + d + c0, opr);
          float d2 = 0.f;
#pragma unroll
          for (int e = 0; e < VW; ++e) {
            const float a = op_round_t<PREC>(sh[e]) - op_round_t<PREC>(osh[e]), b = op_round_t<PREC>(pr[e]) - op_round_t<PREC>(opr[e]);
            d2 = fmaf(a, a, fmaf(b, b, d2));
          }
          sum += sqrtf(warp_sum(d2));
        }
        if (lane == 0) rowsum[t] = sum;
      }
    }
  }
  if (lane == 0) { red[warp][0] = acc_orth; red[warp][1] = acc_ps; red[warp][2] = acc_pp; }
  __syncthreads();
  if (threadIdx.x < 3) {
    float s2 = 0.f;
    for (int w = 0; w < 4; ++w) s2 += red[w][threadIdx.x];
    if (threadIdx.x == 0) s2 /= (float)p.B;
    reinterpret_cast<float*>(ws + p.part1_off)[(size_t)blockIdx.x * 4 + threadIdx.x] = s2;
  }
  if (fuse_intra && live && (p.terms & FOCAL_TERM_TEMPORAL) && p.S > 1 && p.b > 1 && lane < p.nT) {
    const int S = p.S, s = i % S, w0 = warp - s;
    const float* rs = smem_f + (size_t)4 * p.nT * D + 4 * 2 * kMaxT;
    float m = 0.f;
    for (int j = 0; j < S; ++j) m += rs[(w0 + j) * kMaxT + lane];
    for (int r = 0; r < pw.world; ++r)
      reinterpret_cast<float*>(pw.ws[r] + p.mintra_off)[(uint64_t)lane * p.Bpad + i] = m / (float)(S * S - S);
  }
  if (pw.world > 1) peer_epoch_bump(p, ws);        // operands of the owned rows are out
}

// ---------------------------------------------------------------------------------------------------------
// fast finalize: gradient assembly per (row, tensor) -- temporal part, the InfoNCE operands 2t / 2t + 1 of the tensor,
// the orthogonality pairs it takes part in -- with the gradient row in registers until it is stored.  4 rows per block.
//   MAXT == 0: one warp per row walks the tensors in turn (4 warps per block, 16 KB of shared memory at D = 256, M = 2,
//              compiled for 8 resident blocks per SM: the reductions are latency chains and want many warps);
//   MAXT > 0:  one warp per (row, tensor), 4 * nT warps per block: for launches that are a single wave of blocks (row
//              shards), where the dependent chain of one warp, not the issue rate, sets the time.
// Both use the same per-tensor body in the same order, so they produce identical bits.
// ---------------------------------------------------------------------------------------------------------
#ifndef FB_FIN_MINB
#define FB_FIN_MINB 8            // resident blocks per SM the row-walk variant is compiled for (measured: 4 / 6 / 8 / 10
#endif                           // blocks -> 107 / 94 / 79 / 92 us at 8192 rows; 8 = 64 registers, small spill)
template <int VW, int MAXT, int PREC>
__global__ void __launch_bounds__(MAXT > 0 ? 128 * MAXT : 128, MAXT > 0 ? 1 : (VW > 4 ? 4 : FB_FIN_MINB))
finalize_rt_kernel(const __grid_constant__ Plan p,
                                                           const __grid_constant__ FeatPtrs f,
                                                           const __grid_constant__ GradPtrs g,
                                                           const uint8_t* __restrict__ ws) {
  extern __shared__ float smem_f[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nT = p.nT, D = p.D, d = p.d, S = p.S;
  const int r = MAXT > 0 ? warp / nT : warp;         // row within the block
  const int t_begin = MAXT > 0 ? warp - r * nT : 0, t_end = MAXT > 0 ? t_begin + 1 : nT;      // tensors of this warp
  const int i = p.seq0 * S + blockIdx.x * 4 + r;
  const bool live = i < p.seq1 * S;
  float* xs = smem_f + (size_t)r * nT * D;           // the nT tensors of this row
  float* nrm = smem_f + (size_t)4 * nT * D + r * 2 * kMaxT;
  const int c0 = VW * lane;
  if (live) {
    const size_t roff = feat_row_off(p, i);
    for (int t = t_begin; t < t_end; ++t) {
      float sh[VW], pr[VW];
      const float* src = feat_base(p, f, ws, t) + roff;
      ld_frag<VW>(src + c0, sh);
      ld_frag<VW>(src + d + c0, pr);
      st_frag<VW>(xs + t * D + c0, sh);
      st_frag<VW>(xs + t * D + d + c0, pr);
      float a = 0.f, b = 0.f;
#pragma unroll
      for (int e = 0; e < VW; ++e) { a = fmaf(sh[e], sh[e], a); b = fmaf(pr[e], pr[e], b); }
      a = warp_sum(a); b = warp_sum(b);
      if (lane == 0) { nrm[2 * t] = a; nrm[2 * t + 1] = b; }
    }
  }
  __syncthreads();
  if (!live) return;
  const int I = i / S, s = i % S;
  const uint64_t rowN = (uint64_t)s * p.bpad + I;
#pragma unroll 1
  for (int t = t_begin; t < t_end; ++t) {
  float sh[VW], pr[VW];
  ld_frag<VW>(xs + t * D + c0, sh);
  ld_frag<VW>(xs + t * D + d + c0, pr);
  float gsh[VW], gpr[VW];
#pragma unroll
  for (int e = 0; e < VW; ++e) { gsh[e] = 0.f; gpr[e] = 0.f; }

  // ---- temporal part
  if ((p.terms & FOCAL_TERM_TEMPORAL) && p.b > 1 && S > 1) {
    const float bb = (float)p.b * (float)(p.b - 1);
    const int Dp = p.kbFull * p.epb;
    const int extra = __ldg(reinterpret_cast<const int32_t*>(ws + p.flag_tmp_off) + (uint64_t)t * (p.Bpad / kTileM) + i / kTileM);
    float rho = __ldg(reinterpret_cast<const float*>(ws + p.rho_off) + (uint64_t)t * p.Bpad + i);
    const float* y = reinterpret_cast<const float*>(ws + p.dx_off) + ((uint64_t)t * p.Bpad + i) * Dp;
    int cnt = __ldg(reinterpret_cast<const int32_t*>(ws + p.cnt_off) + (uint64_t)t * p.bpad + I);
    float ysh[VW], ypr[VW];
    ld_frag<VW>(y + c0, ysh);
    ld_frag<VW>(y + d + c0, ypr);
    for (int k = 1; k <= extra; ++k) {
      const float* y2 = reinterpret_cast<const float*>(ws + p.dx_off + k * p.dx2_delta) + ((uint64_t)t * p.Bpad + i) * Dp;
      float zsh[VW], zpr[VW];
      ld_frag<VW>(y2 + c0, zsh);
      ld_frag<VW>(y2 + d + c0, zpr);
#pragma unroll
      for (int e = 0; e < VW; ++e) { ysh[e] += zsh[e]; ypr[e] += zpr[e]; }
      rho += __ldg(reinterpret_cast<const float*>(ws + p.rho_off + k * p.rho2_delta) + (uint64_t)t * p.Bpad + i);
      cnt += __ldg(reinterpret_cast<const int32_t*>(ws + p.cnt_off + k * p.cnt2_delta) + (uint64_t)t * p.bpad + I);
    }
    float rsh[VW], rpr[VW];
#pragma unroll
    for (int e = 0; e < VW; ++e) {
      rsh[e] = op_round_t<PREC>(sh[e]); rpr[e] = op_round_t<PREC>(pr[e]);
      gsh[e] = p.w_rank * (rsh[e] * rho - ysh[e]);
      gpr[e] = p.w_rank * (rpr[e] * rho - ypr[e]);
    }
    // intra-sequence pairs: dL/dm_II = cnt / (b(b-1)), spread over S^2 - S ordered pairs, both orders
    const float coef = p.w_rank * 2.f * (float)cnt / (bb * (float)(S * S - S));
    if (cnt > 0) {
      const int r0 = r - s;                              // block row holding position 0 of this sequence
      for (int j = 0; j < S; ++j) {
        if (j == s) continue;
        const float* xo = smem_f + (size_t)(r0 + j) * nT * D + t * D;
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
        if (d2 > 0.f) {
          const float rr = coef * rsqrtf(d2);
#pragma unroll
          for (int e = 0; e < VW; ++e) { gsh[e] = fmaf(rr, osh[e], gsh[e]); gpr[e] = fmaf(rr, opr[e], gpr[e]); }
        }
      }
    }
  }

  // ---- InfoNCE (standard topology: operands 2t = shared half, 2t + 1 = private half of tensor t)
  if (p.terms & FOCAL_TERM_NCE) {
    const float inv_tsn = 1.f / (p.T * (float)S * (float)(2 * p.b));
    const float inv_alpha = 1.f / p.alpha;
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      const OpDesc& op = p.ops[2 * t + half];
      if (op.nuse == 0) continue;
      const int wp = op.kb * p.epb;
      const float inv_nk = fminf(rsqrtf(nrm[2 * t + half]), 1.f / kNceEps);        // 1 / max(|z|, eps)
      const float fk = p.alpha * inv_nk;
      float x[VW], tmp[VW];
#pragma unroll
      for (int e = 0; e < VW; ++e) { x[e] = half ? pr[e] : sh[e]; tmp[e] = 0.f; }
      for (int u = 0; u < op.nuse; ++u) {
        const int q = op.use_prob[u], side = op.use_side[u];
        const ProbDesc& prb = p.probs[q];
        const OpDesc& po = p.ops[op.use_partner[u]];                  // partner operand: positive row p(k)
        const float fp = p.alpha * fminf(rsqrtf(nrm[2 * po.tensor + (po.col0 ? 1 : 0)]), 1.f / kNceEps);
        float px[VW], acc[VW];
        ld_frag<VW>(xs + po.tensor * D + po.col0 + c0, px);
        ld_frag<VW>(reinterpret_cast<const float*>(ws + prb.dz_off) + ((uint64_t)side * S * p.bpad + rowN) * wp + c0, acc);
        const int extra = __ldg(reinterpret_cast<const int32_t*>(ws + p.flag_nce_off) +
                                (((uint64_t)q * S + s) * 2 + side) * (p.bpad / kTileM) + I / kTileM);
        for (int k = 1; k <= extra; ++k) {
          float acc2[VW];
          ld_frag<VW>(reinterpret_cast<const float*>(ws + prb.dz_off + k * p.dz2_delta) + ((uint64_t)side * S * p.bpad + rowN) * wp + c0, acc2);
#pragma unroll
          for (int e = 0; e < VW; ++e) acc[e] += acc2[e];
        }
        const float* rs = reinterpret_cast<const float*>(ws + p.rsum_off) + ((uint64_t)(q * S + s) * 2) * p.bpad;
        const float r_k = __ldg(rs + (uint64_t)side * p.bpad + I), r_p = __ldg(rs + (uint64_t)(1 - side) * p.bpad + I);
        // positive column in fp32 (masked out of the tiles): W_kp - 2 is a tiny difference when the positive dominates
        const float gpos = __ldg(reinterpret_cast<const float*>(ws + p.pos_off) + ((uint64_t)(q * S + s) * 2 + side) * p.bpad + I);
#pragma unroll
        for (int e = 0; e < VW; ++e) px[e] = op_round_t<PREC>(px[e] * fp);
        const float wkp = ex2_approx(gpos) * (__frcp_rn(r_k) + __frcp_rn(r_p));
        const float wq = prb.weight * inv_tsn * inv_alpha;
#pragma unroll
        for (int e = 0; e < VW; ++e) tmp[e] = fmaf(wq, fmaf(wkp - 2.f, px[e], acc[e]), tmp[e]);
      }
      float dot = 0.f;                                // d zh / d z = (I - zh zh^T) / n
#pragma unroll
      for (int e = 0; e < VW; ++e) dot = fmaf(tmp[e], x[e], dot);
      dot = warp_sum(dot) * inv_nk * inv_nk;
#pragma unroll
      for (int e = 0; e < VW; ++e) {
        const float v = fmaf(fmaf(-dot, x[e], tmp[e]), inv_nk, half ? gpr[e] : gsh[e]);
        if (half) gpr[e] = v; else gsh[e] = v;
      }
    }
  }

  // ---- orthogonality: every pair this tensor takes part in (both sides when it pairs its own halves)
  if (p.terms & FOCAL_TERM_ORTH) {
    for (int k = 0; k < p.nOrth; ++k) {
      const OrthDesc& od = p.orth[k];
      if (od.tu != t && od.tv != t) continue;
      float u[VW], v[VW];
      ld_frag<VW>(xs + od.tu * D + od.cu + c0, u);
      ld_frag<VW>(xs + od.tv * D + od.cv + c0, v);
      float dot = 0.f;
#pragma unroll
      for (int e = 0; e < VW; ++e) dot = fmaf(u[e], v[e], dot);
      dot = warp_sum(dot);
      const float nu = nrm[2 * od.tu + (od.cu ? 1 : 0)] + kOrthEps;
      const float nv = nrm[2 * od.tv + (od.cv ? 1 : 0)] + kOrthEps;
      const float inv_den = rsqrtf(nu * nv);
      const float cs = dot * inv_den;
      if (cs >= 0.f) {                                    // clamp_min passes gradient at equality
        const float a = p.w_orth / (float)p.B;
        const float ad = a * inv_den, au = -a * cs * __frcp_rn(nu), av = -a * cs * __frcp_rn(nv);
        if (od.tu == t) {
#pragma unroll
          for (int e = 0; e < VW; ++e) {
            const float w = fmaf(ad, v[e], fmaf(au, u[e], od.cu ? gpr[e] : gsh[e]));
            if (od.cu) gpr[e] = w; else gsh[e] = w;
          }
        }
        if (od.tv == t) {
#pragma unroll
          for (int e = 0; e < VW; ++e) {
            const float w = fmaf(ad, u[e], fmaf(av, v[e], od.cv ? gpr[e] : gsh[e]));
            if (od.cv) gpr[e] = w; else gsh[e] = w;
          }
        }
      }
    }
  }
  float* out = grad_base(p, g, ws, t) + (size_t)i * D;
  st_frag<VW>(out + c0, gsh);
  st_frag<VW>(out + d + c0, gpr);
  }  // tensors of this warp
}

}  // namespace fb
