// Row kernels, third generation: ONE WARP PER (sequence, tensor), a lane group per row.
//
// The first two generations (row_kernels_fast.cuh: a warp per row walking the tensors; row_kernels_v2.cuh: a warp per
// (row, tensor)) were bound by instruction issue, not by HBM (ncu: 746 / 1355 warp instructions per (row, tensor),
// issue slots 62-68 % busy, 29-30 % of the measured HBM bandwidth): a warp only ever owned 8 elements per lane, so the
// per-row overhead -- 5-step shuffle reductions, swizzled 64-bit address arithmetic, plan decoding, staging rows in shared
// memory for the neighbours -- outweighed the element work several times, and 32 bytes in flight per lane could not
// cover the HBM latency.
//
// Here a warp owns the S rows of one sequence of one tensor (S in {1, 2, 4}): lane = g * LPR + l, LPR = 32 / S lanes
// per row, row g of the sequence.  Lane l owns the float4 slots k = 0..NQ-1 at columns 4 * (k * LPR + l) of the shared
// half and the same columns of the private half (NQ = d * S / 128: 4 for D = 256, S = 4), so
//   * every global access is a fully coalesced 16-byte (features, accumulators, gradients) or 8-byte (bf16 operands) vector
//     and a lane keeps 2 * NQ independent 16-byte loads in flight (128 B at the headline shape);
//   * per-row reductions are log2(LPR) shuffle steps that serve S rows at once, and they are batched;
//   * the other rows of the sequence (intra-sequence distances m_II and their gradient) are one __shfl_xor away -- no
//     shared-memory staging, no second pass over the row;
//   * the prologue keeps the squared pair distances it computes anyway (pd), finalize reads them instead of
//     re-reducing; finalize takes the positive-pair rows from the bf16 operand arrays the prologue wrote (exactly the
//     values the tiles saw) instead of re-normalising and re-rounding the partner's fp32 row.
// A block holds SEQB = 8 / nT consecutive sequences x all nT tensors (nT <= 8); the only block-wide exchange is the private
// halves + norms of the other modalities of the view for the cross-modality orthogonality pairs (shared memory).
// Shapes off this path (odd widths, noPrivate, S not in {1, 2, 4}, D = 512 with S = 4) keep using the older kernels.
#pragma once
#include "peer.cuh"
#include "plan.h"
#include "ptx.cuh"
#include "row_kernels.cuh"

namespace fb {

__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ float f4_get(const float4& v, int e) { return e == 0 ? v.x : (e == 1 ? v.y : (e == 2 ? v.z : v.w)); }

// Blocks per SM ahead whose feature rows a block prefetches into L2 (0 = off).  Measured (profiles/r2_variants.txt): 0 / 2 /
// 4 ahead = 44.8 / 44.6 / 44.7 us prologue, 65.3 / 67.0 / 67.2 us finalize: the first-phase latency is not what bounds
// either kernel, so it stays off.
#ifndef FB_L2_HINTS
#define FB_L2_HINTS 0           // see gram_kernel.cuh
#endif
#ifndef FB_ROW_PF_AHEAD
#define FB_ROW_PF_AHEAD 0
#endif
// Start moving [p, p + bytes) (16-byte aligned, a multiple of 16) towards L2: no register, no scoreboard.
__device__ __forceinline__ void prefetch_l2(const void* p, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}

// reduce N values over the LPR lanes of a row group (all lanes of the group end up with the sums); butterflies interleaved
template <int LPR, int N>
__device__ __forceinline__ void group_sum_n(float (&v)[N]) {
#pragma unroll
  for (int o = LPR / 2; o > 0; o >>= 1) {
#pragma unroll
    for (int k = 0; k < N; ++k) v[k] += __shfl_xor_sync(0xffffffffu, v[k], o);
  }
}

// the value a tile sees for operand element v, and the bf16 images it is stored as: hi (bf16 mode: the only image), lo
template <int PREC>
__device__ __forceinline__ void round4(const float4& x, float4& r, uint2& hi, uint2& lo) {
  hi.x = pack_bf16x2(x.x, x.y);
  hi.y = pack_bf16x2(x.z, x.w);
  r.x = __uint_as_float(hi.x << 16); r.y = __uint_as_float(hi.x & 0xffff0000u);
  r.z = __uint_as_float(hi.y << 16); r.w = __uint_as_float(hi.y & 0xffff0000u);
  if (PREC == FOCAL_PREC_FP32) {
    lo.x = pack_bf16x2(x.x - r.x, x.y - r.y);
    lo.y = pack_bf16x2(x.z - r.z, x.w - r.w);
    r.x += __uint_as_float(lo.x << 16); r.y += __uint_as_float(lo.x & 0xffff0000u);
    r.z += __uint_as_float(lo.y << 16); r.w += __uint_as_float(lo.y & 0xffff0000u);
  } else {
    lo = make_uint2(0u, 0u);
  }
}
// the fp32 value of 4 stored operand elements (hi image, plus the lo image for split tiles)
template <int PREC>
__device__ __forceinline__ float4 unround4(const uint2& hi, const uint2& lo) {
  float4 r = make_float4(__uint_as_float(hi.x << 16), __uint_as_float(hi.x & 0xffff0000u), __uint_as_float(hi.y << 16),
                         __uint_as_float(hi.y & 0xffff0000u));
  if (PREC == FOCAL_PREC_FP32) {
    r.x += __uint_as_float(lo.x << 16); r.y += __uint_as_float(lo.x & 0xffff0000u);
    r.z += __uint_as_float(lo.y << 16); r.w += __uint_as_float(lo.y & 0xffff0000u);
  }
  return r;
}
// sum over 4 stored operand elements of tile_sq (row_kernels.cuh): bf16 tiles hi^2, split tiles (hi + lo)^2
template <int PREC>
__device__ __forceinline__ float tile_sq4(const uint2& hi, const uint2& lo, float acc) {
  const float4 r = unround4<PREC>(hi, lo);
  return fmaf(r.x, r.x, fmaf(r.y, r.y, fmaf(r.z, r.z, fmaf(r.w, r.w, acc))));
}

// Lanes 2j and 2j + 1 of a row hold the two 8-byte halves of one 16-byte operand chunk, for slot k and for slot k + 1.
// After one exchange the even lane holds the whole chunk of slot k (`a`) and the odd lane the whole chunk of slot k + 1
// (`b`): every store instruction then writes full 128-byte operand rows in 16-byte pieces -- half the store
// instructions, and the granularity remote (NVLink) writes want.
__device__ __forceinline__ uint4 pair_chunk(const uint2& a, const uint2& b, bool odd) {
  const uint2 send = odd ? a : b;
  const uint32_t gx = __shfl_xor_sync(0xffffffffu, send.x, 1), gy = __shfl_xor_sync(0xffffffffu, send.y, 1);
  return odd ? make_uint4(gx, gy, b.x, b.y) : make_uint4(a.x, a.y, gx, gy);
}
// byte offset of the 16-byte chunk that holds elements [c, c + 8) (c % 8 == 0) of row `row`
__device__ __forceinline__ uint64_t op_off16(uint64_t krows, uint64_t row, int c) {
  return ((uint64_t)(c >> 6) * krows + row) * 128 + ((((uint32_t)(c & 63) >> 3) ^ (uint32_t)(row & 7)) << 4);
}
// byte offset of the 8-byte group that holds elements [c, c + 4) (c % 4 == 0) of row `row` in a swizzled operand array
// with `krows` rows per K block
__device__ __forceinline__ uint64_t op_off8(uint64_t krows, uint64_t row, int c) {
  return ((uint64_t)(c >> 6) * krows + row) * 128 + ((((uint32_t)(c & 63) >> 3) ^ (uint32_t)(row & 7)) << 4) + (c & 4) * 2;
}

// ---------------------------------------------------------------------------------------------------------
// prologue: norms, InfoNCE + temporal operands, orthogonality terms, intra-sequence mean distances m_II
// ---------------------------------------------------------------------------------------------------------
template <int S, int NQ, int PREC>
__global__ void __launch_bounds__(256, 2) prologue_v3_kernel(const __grid_constant__ Plan p,
                                                             const __grid_constant__ FeatPtrs f,
                                                             const __grid_constant__ PeerWs pw,
                                                             uint8_t* __restrict__ ws) {
  constexpr int LPR = 32 / S;
  extern __shared__ float smem_f[];
  pdl_launch_dependents();
  pdl_wait();
  // blocks behind the row blocks (single-GPU launches only): zero the padding rows of the operand arrays -- no launch of
  // its own for that
  if ((int)blockIdx.x >= p.nblk1) {
    zero_pad_body(p, ws, (long)(blockIdx.x - p.nblk1) * blockDim.x + threadIdx.x, (long)(gridDim.x - p.nblk1) * blockDim.x);
    return;
  }
  const int nT = p.nT, d = p.d, seqb = p.seqb;
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;   // warp-uniform
  const int q = warp / nT, t = warp - q * nT;                 // sequence within the block, tensor
  const int g = lane / LPR, l = lane % LPR;                   // row of the sequence, lane within the row
  // Row-sharded launches may be replicated gridDim.y times (FOCAL_B200_PROLOGUE_REPLICAS; default 1): replica j computes the
  // same rows and stores them into the workspaces of the ranks r with r % gridDim.y == j.  Measured SLOWER than one
  // replica (79 vs 60 us at 4 GPUs, profiles/r2_prologue_dbg_4.txt): the exchange is bound by NVLink ingress, not by
  // the number of warps pushing stores.  With the NVSwitch multicast mapping (pw.mc) replica 0 alone stores.
  const int rep = blockIdx.y, nrep = gridDim.y;
  const bool rep0 = rep == 0;
  const int I0 = p.local_rows ? p.seq0 : 0, I1 = p.local_rows ? p.seq1 : p.b;
  const int I = I0 + blockIdx.x * seqb + q;
  const bool live = I < I1;
  const bool owned = live && I >= p.seq0 && I < p.seq1;
  const int i = I * S + g;                                    // feature row
  float* prv = smem_f;                                        // [seqb][nT][S][d] private halves (cross-modality orth pairs)
  float* nbs = prv + (size_t)seqb * nT * S * d;               // [seqb][nT][S] squared norm of the private half
  float* red = nbs + seqb * nT * S;                           // [seqb * nT] orthogonality partial sums
  const bool tmp_on = (p.terms & FOCAL_TERM_TEMPORAL) != 0;
  const bool orth_on = (p.terms & FOCAL_TERM_ORTH) != 0;
  float4 sh[NQ], pr[NQ];
  float nb = 1.f, acc_orth = 0.f;
  if (live) {
    const float* src = feat_base(p, f, ws, t) + feat_row_off(p, i);
#pragma unroll
    for (int k = 0; k < NQ; ++k) sh[k] = (FB_L2_HINTS & 2) ? ld4_nc_hint(src + 4 * (k * LPR + l), l2_policy_evict_first()) : ldg4(src + 4 * (k * LPR + l));
#pragma unroll
    for (int k = 0; k < NQ; ++k) pr[k] = (FB_L2_HINTS & 2) ? ld4_nc_hint(src + d + 4 * (k * LPR + l), l2_policy_evict_first()) : ldg4(src + d + 4 * (k * LPR + l));
    // the (sequence, tensor) that the block taking this SM slot next will read: start it towards L2 now
    if (FB_ROW_PF_AHEAD) {
      const int In = I + FB_ROW_PF_AHEAD * p.num_sms * seqb;
      if (In < I1 && lane == 0) prefetch_l2(feat_base(p, f, ws, t) + feat_row_off(p, In * S), (uint32_t)(S * p.D * 4));
    }
    // ---- rounded rows (what the temporal tiles see), norms, shared . private
    uint2 hsh[NQ], lsh[NQ], hpr[NQ], lpr[NQ];
    float q4[4] = {0.f, 0.f, 0.f, 0.f};                       // |shared|^2, |private|^2, |rounded row|^2 (tile product), shared . private
#pragma unroll
    for (int k = 0; k < NQ; ++k) {
      float4 rnd;
      round4<PREC>(sh[k], rnd, hsh[k], lsh[k]);
      round4<PREC>(pr[k], rnd, hpr[k], lpr[k]);
      q4[0] = fmaf(sh[k].x, sh[k].x, fmaf(sh[k].y, sh[k].y, fmaf(sh[k].z, sh[k].z, fmaf(sh[k].w, sh[k].w, q4[0]))));
      q4[1] = fmaf(pr[k].x, pr[k].x, fmaf(pr[k].y, pr[k].y, fmaf(pr[k].z, pr[k].z, fmaf(pr[k].w, pr[k].w, q4[1]))));
      q4[3] = fmaf(sh[k].x, pr[k].x, fmaf(sh[k].y, pr[k].y, fmaf(sh[k].z, pr[k].z, fmaf(sh[k].w, pr[k].w, q4[3]))));
      q4[2] = tile_sq4<PREC>(hsh[k], lsh[k], q4[2]);
      q4[2] = tile_sq4<PREC>(hpr[k], lpr[k], q4[2]);
    }
    group_sum_n<LPR, 4>(q4);
    const float na = q4[0];
    nb = q4[1];
    if (l == 0) {
      if (rep0) *reinterpret_cast<float2*>(ws + p.nrm_off + ((uint64_t)t * p.Bpad + i) * 8) = make_float2(na, nb);
      nbs[(q * nT + t) * S + g] = nb;
    }
    if (orth_on && p.M > 1) {
      float* mine = prv + ((size_t)(q * nT + t) * S + g) * d;
#pragma unroll
      for (int k = 0; k < NQ; ++k) *reinterpret_cast<float4*>(mine + 4 * (k * LPR + l)) = pr[k];
    }
    // ---- InfoNCE operands: x / max(|x|, eps) * sqrt(log2 e / T), position-major rows
    if (p.terms & FOCAL_TERM_NCE) {
      const float fa = p.alpha * fminf(rsqrtf(na), 1.f / kNceEps), fb2 = p.alpha * fminf(rsqrtf(nb), 1.f / kNceEps);
      const uint64_t rowN = (uint64_t)g * p.bpad + I, rowsNce = (uint64_t)S * p.bpad;
      const uint64_t off_s = p.ops[2 * t].off, off_p = p.ops[2 * t + 1].off;
      const uint64_t lo_img = (uint64_t)(p.ops[2 * t].kb / 2) * rowsNce * 128;      // split tiles: lo image behind the hi blocks
      const bool odd = (l & 1) != 0;
#pragma unroll
      for (int k = 0; k + 1 < NQ; k += 2) {                   // slot pairs: 16-byte chunks after a lane-pair exchange
        uint2 hs[2], ls[2], hp[2], lp[2];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          const float4 zs = make_float4(sh[k + u].x * fa, sh[k + u].y * fa, sh[k + u].z * fa, sh[k + u].w * fa);
          const float4 zp = make_float4(pr[k + u].x * fb2, pr[k + u].y * fb2, pr[k + u].z * fb2, pr[k + u].w * fb2);
          float4 rr;
          round4<PREC>(zs, rr, hs[u], ls[u]);
          round4<PREC>(zp, rr, hp[u], lp[u]);
        }
        const uint4 cs = pair_chunk(hs[0], hs[1], odd), cp = pair_chunk(hp[0], hp[1], odd);
        uint4 cls = make_uint4(0u, 0u, 0u, 0u), clp = cls;
        if (PREC == FOCAL_PREC_FP32) { cls = pair_chunk(ls[0], ls[1], odd); clp = pair_chunk(lp[0], lp[1], odd); }
        const uint64_t o = op_off16(rowsNce, rowN, 4 * ((k + (odd ? 1 : 0)) * LPR + (l & ~1)));
        if (pw.mc) {                                          // one multicast store serves every rank
          if (rep0) {
            uint8_t* w = pw.mc;
            mc_st16(w + off_s + o, cs);
            mc_st16(w + off_p + o, cp);
            if (PREC == FOCAL_PREC_FP32) {
              mc_st16(w + off_s + lo_img + o, cls);
              mc_st16(w + off_p + lo_img + o, clp);
            }
          }
        } else for (int rk = rep; rk < pw.world; rk += nrep) {
          uint8_t* w = pw.ws[rk];
          *reinterpret_cast<uint4*>(w + off_s + o) = cs;
          *reinterpret_cast<uint4*>(w + off_p + o) = cp;
          if (PREC == FOCAL_PREC_FP32) {
            *reinterpret_cast<uint4*>(w + off_s + lo_img + o) = cls;
            *reinterpret_cast<uint4*>(w + off_p + lo_img + o) = clp;
          }
        }
      }
      if (NQ & 1) {                                           // odd slot count: the last slot goes out in 8-byte pieces
        constexpr int k = NQ - 1;
        const int c = 4 * (k * LPR + l);
        const float4 zs = make_float4(sh[k].x * fa, sh[k].y * fa, sh[k].z * fa, sh[k].w * fa);
        const float4 zp = make_float4(pr[k].x * fb2, pr[k].y * fb2, pr[k].z * fb2, pr[k].w * fb2);
        float4 rr;
        uint2 hs, ls, hp, lp;
        round4<PREC>(zs, rr, hs, ls);
        round4<PREC>(zp, rr, hp, lp);
        const uint64_t o = op_off8(rowsNce, rowN, c);
        if (pw.mc) {                                          // one multicast store serves every rank
          if (rep0) {
            uint8_t* w = pw.mc;
            mc_st8(w + off_s + o, hs);
            mc_st8(w + off_p + o, hp);
            if (PREC == FOCAL_PREC_FP32) {
              mc_st8(w + off_s + lo_img + o, ls);
              mc_st8(w + off_p + lo_img + o, lp);
            }
          }
        } else for (int rk = rep; rk < pw.world; rk += nrep) {
          uint8_t* w = pw.ws[rk];
          *reinterpret_cast<uint2*>(w + off_s + o) = hs;
          *reinterpret_cast<uint2*>(w + off_p + o) = hp;
          if (PREC == FOCAL_PREC_FP32) {
            *reinterpret_cast<uint2*>(w + off_s + lo_img + o) = ls;
            *reinterpret_cast<uint2*>(w + off_p + lo_img + o) = lp;
          }
        }
      }
      if (d & 63) {                                           // d = 32 or 96: zero the unused half of the last K block
        const uint64_t o = op_off8(rowsNce, rowN, 4 * (NQ * LPR + l));
        if (pw.mc) {                                          // one multicast store serves every rank
          if (rep0) {
            uint8_t* w = pw.mc;
            mc_st8(w + off_s + o, make_uint2(0u, 0u));
            mc_st8(w + off_p + o, make_uint2(0u, 0u));
            if (PREC == FOCAL_PREC_FP32) {
              mc_st8(w + off_s + lo_img + o, make_uint2(0u, 0u));
              mc_st8(w + off_p + lo_img + o, make_uint2(0u, 0u));
            }
          }
        } else for (int rk = rep; rk < pw.world; rk += nrep) {
          uint8_t* w = pw.ws[rk];
          *reinterpret_cast<uint2*>(w + off_s + o) = make_uint2(0u, 0u);
          *reinterpret_cast<uint2*>(w + off_p + o) = make_uint2(0u, 0u);
          if (PREC == FOCAL_PREC_FP32) {
            *reinterpret_cast<uint2*>(w + off_s + lo_img + o) = make_uint2(0u, 0u);
            *reinterpret_cast<uint2*>(w + off_p + lo_img + o) = make_uint2(0u, 0u);
          }
        }
      }
    }
    // ---- temporal operands (raw rows, natural order) + squared norm of what the tiles will see
    if (tmp_on) {
      const uint64_t krows = (uint64_t)p.Bpad;
      const uint64_t xoff = p.xt_off + (uint64_t)t * p.kbFull * krows * 128;
      const uint64_t lo_img = (uint64_t)(p.kbFull / 2) * krows * 128;
      const bool odd = (l & 1) != 0;
#pragma unroll
      for (int k = 0; k + 1 < NQ; k += 2) {                   // slot pairs -> 16-byte chunks, full 128-byte rows per store
        const uint4 cs = pair_chunk(hsh[k], hsh[k + 1], odd), cp = pair_chunk(hpr[k], hpr[k + 1], odd);
        uint4 cls = make_uint4(0u, 0u, 0u, 0u), clp = cls;
        if (PREC == FOCAL_PREC_FP32) { cls = pair_chunk(lsh[k], lsh[k + 1], odd); clp = pair_chunk(lpr[k], lpr[k + 1], odd); }
        const int cc = 4 * ((k + (odd ? 1 : 0)) * LPR + (l & ~1));
        const uint64_t o1 = op_off16(krows, (uint64_t)i, cc), o2 = op_off16(krows, (uint64_t)i, d + cc);
        if (pw.mc) {                                          // one multicast store serves every rank
          if (rep0) {
            uint8_t* xt = pw.mc + xoff;
            mc_st16(xt + o1, cs);
            mc_st16(xt + o2, cp);
            if (PREC == FOCAL_PREC_FP32) {
              mc_st16(xt + lo_img + o1, cls);
              mc_st16(xt + lo_img + o2, clp);
            }
          }
        } else for (int rk = rep; rk < pw.world; rk += nrep) {
          uint8_t* xt = pw.ws[rk] + xoff;
          *reinterpret_cast<uint4*>(xt + o1) = cs;
          *reinterpret_cast<uint4*>(xt + o2) = cp;
          if (PREC == FOCAL_PREC_FP32) {
            *reinterpret_cast<uint4*>(xt + lo_img + o1) = cls;
            *reinterpret_cast<uint4*>(xt + lo_img + o2) = clp;
          }
        }
      }
      if (NQ & 1) {
        constexpr int k = NQ - 1;
        const int c = 4 * (k * LPR + l);
        const uint64_t o1 = op_off8(krows, (uint64_t)i, c), o2 = op_off8(krows, (uint64_t)i, d + c);
        if (pw.mc) {                                          // one multicast store serves every rank
          if (rep0) {
            uint8_t* xt = pw.mc + xoff;
            mc_st8(xt + o1, hsh[k]);
            mc_st8(xt + o2, hpr[k]);
            if (PREC == FOCAL_PREC_FP32) {
              mc_st8(xt + lo_img + o1, lsh[k]);
              mc_st8(xt + lo_img + o2, lpr[k]);
            }
          }
        } else for (int rk = rep; rk < pw.world; rk += nrep) {
          uint8_t* xt = pw.ws[rk] + xoff;
          *reinterpret_cast<uint2*>(xt + o1) = hsh[k];
          *reinterpret_cast<uint2*>(xt + o2) = hpr[k];
          if (PREC == FOCAL_PREC_FP32) {
            *reinterpret_cast<uint2*>(xt + lo_img + o1) = lsh[k];
            *reinterpret_cast<uint2*>(xt + lo_img + o2) = lpr[k];
          }
        }
      }
      if (pw.mc) {
        if (l == 0 && rep0) mc_st4(pw.mc + p.sq_off + ((uint64_t)t * p.Bpad + i) * 4, q4[2]);
      } else if (l < pw.world && l % nrep == rep) {
        reinterpret_cast<float*>(pw.ws[l] + p.sq_off)[(uint64_t)t * p.Bpad + i] = q4[2];
      }
    }
    // ---- orthogonality (loss.py:96-104), pair (shared_t, private_t)
    if (owned && orth_on) acc_orth = fmaxf(q4[3] * rsqrtf((na + kOrthEps) * (nb + kOrthEps)), 0.f);
    // ---- intra-sequence distances of the rounded rows: the partner rows live in the other lane groups of this warp.
    // Row g takes its distances to rows g ^ 1 .. g ^ (S-1); both ends of a pair compute bit-identical sums.
    if (S > 1 && tmp_on && p.b > 1) {
      float d2[S > 1 ? S - 1 : 1];
#pragma unroll
      for (int j = 1; j < S; ++j) {
        float a = 0.f;
#pragma unroll
        for (int k = 0; k < NQ; ++k) {
          const float4 rs = unround4<PREC>(hsh[k], lsh[k]), rp = unround4<PREC>(hpr[k], lpr[k]);
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float u = f4_get(rs, e) - __shfl_xor_sync(0xffffffffu, f4_get(rs, e), j * LPR);
            const float v = f4_get(rp, e) - __shfl_xor_sync(0xffffffffu, f4_get(rp, e), j * LPR);
            a = fmaf(u, u, fmaf(v, v, a));
          }
        }
        d2[j - 1] = a;
      }
      group_sum_n<LPR, (S > 1 ? S - 1 : 1)>(d2);
      float m = 0.f;
#pragma unroll
      for (int j = 1; j < S; ++j) m += sqrtf(d2[j - 1]);
#pragma unroll
      for (int o = LPR; o < 32; o <<= 1) m += __shfl_xor_sync(0xffffffffu, m, o);       // ordered pairs of the sequence
      m = m / (float)(S * S - S);
      if (l == 0 && rep0) {
        float4 pd4 = make_float4(d2[0], S > 2 ? d2[(S > 2) ? 1 : 0] : 0.f, S > 2 ? d2[(S > 2) ? 2 : 0] : 0.f, 0.f);
        *reinterpret_cast<float4*>(ws + p.pd_off + ((uint64_t)t * p.Bpad + i) * 16) = pd4;
      }
      if (pw.mc) {
        if (l == 0 && rep0) mc_st4(pw.mc + p.mintra_off + ((uint64_t)t * p.Bpad + i) * 4, m);
      } else if (l < pw.world && l % nrep == rep) {
        reinterpret_cast<float*>(pw.ws[l] + p.mintra_off)[(uint64_t)t * p.Bpad + i] = m;
      }
    }
  }
  __syncthreads();
  // ---- orthogonality, pairs (private_t, private_t') with t' > t of the same view
  if (owned && orth_on) {
    const int vend = (t / p.M + 1) * p.M;
    for (int t2 = t + 1; t2 < vend; ++t2) {
      const float* other = prv + ((size_t)(q * nT + t2) * S + g) * d;
      float dot[1] = {0.f};
#pragma unroll
      for (int k = 0; k < NQ; ++k) {
        const float4 v = ld4(other + 4 * (k * LPR + l));
        dot[0] = fmaf(pr[k].x, v.x, fmaf(pr[k].y, v.y, fmaf(pr[k].z, v.z, fmaf(pr[k].w, v.w, dot[0]))));
      }
      group_sum_n<LPR, 1>(dot);
      acc_orth += fmaxf(dot[0] * rsqrtf((nb + kOrthEps) * (nbs[(q * nT + t2) * S + g] + kOrthEps)), 0.f);
    }
  }
  // per-warp sum over its rows (fixed order), then a fixed-order block sum: deterministic
  {
    float a = (l == 0) ? acc_orth : 0.f;
#pragma unroll
    for (int o = LPR; o < 32; o <<= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
    if (lane == 0) red[warp] = a;
  }
  __syncthreads();
  if (threadIdx.x == 0 && rep0) {
    float s2 = 0.f;
    for (int w = 0; w < seqb * nT; ++w) s2 += red[w];
    float* p1 = reinterpret_cast<float*>(ws + p.part1_off) + (size_t)blockIdx.x * 4;
    p1[0] = s2 / (float)p.B; p1[1] = 0.f; p1[2] = 0.f;
  }
  if (pw.world > 1) peer_epoch_bump(p, ws);      // operands of the owned rows are out
}

// 4 consecutive bf16 operand elements -> fp32 (split tiles: hi + lo image)
template <int PREC>
__device__ __forceinline__ float4 ld_op4(const uint8_t* op, uint64_t lo_img, uint64_t o) {
  const uint2 h = __ldg(reinterpret_cast<const uint2*>(op + o));
  float4 r = make_float4(__uint_as_float(h.x << 16), __uint_as_float(h.x & 0xffff0000u), __uint_as_float(h.y << 16),
                         __uint_as_float(h.y & 0xffff0000u));
  if (PREC == FOCAL_PREC_FP32) {
    const uint2 lw = __ldg(reinterpret_cast<const uint2*>(op + lo_img + o));
    r.x += __uint_as_float(lw.x << 16); r.y += __uint_as_float(lw.x & 0xffff0000u);
    r.z += __uint_as_float(lw.y << 16); r.w += __uint_as_float(lw.y & 0xffff0000u);
  }
  return r;
}

// ---------------------------------------------------------------------------------------------------------
// finalize: gradient rows of one (sequence, tensor) = temporal part + the InfoNCE operands of the tensor + its
// orthogonality pairs
// ---------------------------------------------------------------------------------------------------------
#ifndef FB_FIN_MINBLOCKS
#define FB_FIN_MINBLOCKS 2
#endif
template <int S, int NQ, int PREC>
__global__ void __launch_bounds__(256, FB_FIN_MINBLOCKS) finalize_v3_kernel(const __grid_constant__ Plan p,
                                                             const __grid_constant__ FeatPtrs f,
                                                             const __grid_constant__ GradPtrs gp,
                                                             const __grid_constant__ PeerWs pw,
                                                             uint8_t* __restrict__ ws, float* __restrict__ loss5,
                                                             int nce_blocks_valid, int temporal_nan) {
  constexpr int LPR = 32 / S;
  extern __shared__ float smem_f[];
  pdl_launch_dependents();
  pdl_wait();
  // The last block adds up the loss partials of the step (they were all written by earlier launches): the step needs no
  // separate loss_reduce launch, and on the row-sharded path the ranks' loss exchange overlaps the gradient rows.
  if (blockIdx.x == gridDim.x - 1) {
    loss_reduce_body(p, pw, ws, loss5, nce_blocks_valid, temporal_nan);
    return;
  }
  const int nT = p.nT, d = p.d, seqb = p.seqb;
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;   // warp-uniform
  const int q = warp / nT, t = warp - q * nT;
  const int g = lane / LPR, l = lane % LPR;
  const int I = p.seq0 + blockIdx.x * seqb + q;
  const bool live = I < p.seq1;
  const int i = I * S + g;
  float* prv = smem_f;                                        // [seqb][nT][S][d] private halves
  const bool orth_on = (p.terms & FOCAL_TERM_ORTH) != 0;
  float4 sh[NQ], pr[NQ];
  float na = 1.f, nb = 1.f;
  if (live) {
    const float* src = feat_base(p, f, ws, t) + feat_row_off(p, i);
#pragma unroll
    for (int k = 0; k < NQ; ++k) sh[k] = (FB_L2_HINTS & 2) ? ld4_nc_hint(src + 4 * (k * LPR + l), l2_policy_evict_first()) : ldg4(src + 4 * (k * LPR + l));
#pragma unroll
    for (int k = 0; k < NQ; ++k) pr[k] = (FB_L2_HINTS & 2) ? ld4_nc_hint(src + d + 4 * (k * LPR + l), l2_policy_evict_first()) : ldg4(src + d + 4 * (k * LPR + l));
    const float2 n2 = __ldg(reinterpret_cast<const float2*>(ws + p.nrm_off + ((uint64_t)t * p.Bpad + i) * 8));
    if (FB_ROW_PF_AHEAD) {        // feature rows of the block that takes this SM slot next
      const int In = I + FB_ROW_PF_AHEAD * p.num_sms * seqb;
      if (In < p.seq1 && lane == 0) prefetch_l2(feat_base(p, f, ws, t) + feat_row_off(p, In * S), (uint32_t)(S * p.D * 4));
    }
    // ---- The phases below read the accumulator rows of this (sequence, tensor) one after the other, and some of those
    // reads depend on flags (how many stream-K pieces a row block has): a chain of DRAM round trips per warp, which is what
    // bounded this kernel (ncu: 25 % issue-active, long-scoreboard stalls, 2.4 TB/s).  So, while the feature rows are in
    // flight, ask for everything else to be brought to L2: the later phases then run at L2 latency.
    if ((p.terms & FOCAL_TERM_TEMPORAL) && p.b > 1 && S > 1) {
      const int Dp = p.kbFull * p.epb;
      const uint64_t t0 = (uint64_t)t * p.Bpad + (uint64_t)I * S;                 // the S rows of the sequence are adjacent
      const int extra = __ldg(reinterpret_cast<const int32_t*>(ws + p.flag_tmp_off) + (uint64_t)t * (p.Bpad / kTileM) + i / kTileM);
      if (lane == 0) {
        prefetch_l2(ws + p.dx_off + t0 * Dp * 4, (uint32_t)(S * Dp * 4));
        for (int e = 1; e <= extra; ++e) prefetch_l2(ws + p.dx_off + e * p.dx2_delta + t0 * Dp * 4, (uint32_t)(S * Dp * 4));
      }
    }
    if (p.terms & FOCAL_TERM_NCE) {
      const uint64_t rowN = (uint64_t)g * p.bpad + I, rowsNce = (uint64_t)S * p.bpad;
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        const OpDesc& op = p.ops[2 * t + half];
        const int wp = op.kb * p.epb;
        for (int u = 0; u < op.nuse; ++u) {
          const int qp = op.use_prob[u], side = op.use_side[u];
          const OpDesc& po = p.ops[op.use_partner[u]];
          const uint64_t arow = ((uint64_t)side * S * p.bpad + rowN) * wp;
          const uint64_t fidx = ((uint64_t)qp * S + g) * 2 + side;
          const int extra = __ldg(reinterpret_cast<const int32_t*>(ws + p.flag_nce_off) + fidx * (p.bpad / kTileM) + I / kTileM);
          if (l == 0) {
            prefetch_l2(ws + p.probs[qp].dz_off + arow * 4, (uint32_t)(wp * 4));
            for (int e = 1; e <= extra; ++e) prefetch_l2(ws + p.probs[qp].dz_off + e * p.dz2_delta + arow * 4, (uint32_t)(wp * 4));
          }
          if (l >= 1 && l <= po.kb) prefetch_l2(ws + po.off + ((uint64_t)(l - 1) * rowsNce + rowN) * 128, 128u);
        }
      }
    }
    na = n2.x; nb = n2.y;
    if (orth_on && p.M > 1) {
      float* mine = prv + ((size_t)(q * nT + t) * S + g) * d;
#pragma unroll
      for (int k = 0; k < NQ; ++k) *reinterpret_cast<float4*>(mine + 4 * (k * LPR + l)) = pr[k];
    }
  }
  __syncthreads();
  if (!live) return;
  float4 gsh[NQ], gpr[NQ];
#pragma unroll
  for (int k = 0; k < NQ; ++k) { gsh[k] = make_float4(0.f, 0.f, 0.f, 0.f); gpr[k] = make_float4(0.f, 0.f, 0.f, 0.f); }

  // ---- temporal part: x~_i rho_i - (R X~)_i, then the exact intra-sequence pairs
  if ((p.terms & FOCAL_TERM_TEMPORAL) && p.b > 1 && S > 1) {
    const int Dp = p.kbFull * p.epb;
    const uint64_t ti = (uint64_t)t * p.Bpad + i;
    const int extra = __ldg(reinterpret_cast<const int32_t*>(ws + p.flag_tmp_off) + (uint64_t)t * (p.Bpad / kTileM) + i / kTileM);
    float rho = __ldg(reinterpret_cast<const float*>(ws + p.rho_off) + ti);
    int cnt = __ldg(reinterpret_cast<const int32_t*>(ws + p.cnt_off) + (uint64_t)t * p.bpad + I);
    const float4 pd4 = __ldg(reinterpret_cast<const float4*>(ws + p.pd_off + ti * 16));
    const float* y = reinterpret_cast<const float*>(ws + p.dx_off) + ti * Dp;
    float4 ysh[NQ], ypr[NQ];
#pragma unroll
    for (int k = 0; k < NQ; ++k) ysh[k] = ldg4(y + 4 * (k * LPR + l));
#pragma unroll
    for (int k = 0; k < NQ; ++k) ypr[k] = ldg4(y + d + 4 * (k * LPR + l));
    for (int e = 1; e <= extra; ++e) {                        // stream-K: accumulator copies of the secondary pieces
      const float* y2 = reinterpret_cast<const float*>(ws + p.dx_off + e * p.dx2_delta) + ti * Dp;
#pragma unroll
      for (int k = 0; k < NQ; ++k) {
        const float4 a = ldg4(y2 + 4 * (k * LPR + l)), c = ldg4(y2 + d + 4 * (k * LPR + l));
        ysh[k].x += a.x; ysh[k].y += a.y; ysh[k].z += a.z; ysh[k].w += a.w;
        ypr[k].x += c.x; ypr[k].y += c.y; ypr[k].z += c.z; ypr[k].w += c.w;
      }
      rho += __ldg(reinterpret_cast<const float*>(ws + p.rho_off + e * p.rho2_delta) + ti);
      cnt += __ldg(reinterpret_cast<const int32_t*>(ws + p.cnt_off + e * p.cnt2_delta) + (uint64_t)t * p.bpad + I);
    }
    // dL/dm_II = cnt / (b (b-1)), spread over S^2 - S ordered pairs, both orders: weight of pair (i, j) = coef / delta_ij
    const float coef = p.w_rank * 2.f * (float)cnt / ((float)p.b * (float)(p.b - 1) * (float)(S * S - S));
    float rr[S > 1 ? S - 1 : 1];
    rr[0] = pd4.x > 0.f ? coef * rsqrtf(pd4.x) : 0.f;
    if (S > 2) { rr[(S > 2) ? 1 : 0] = pd4.y > 0.f ? coef * rsqrtf(pd4.y) : 0.f; rr[(S > 2) ? 2 : 0] = pd4.z > 0.f ? coef * rsqrtf(pd4.z) : 0.f; }
    float rtot = p.w_rank * rho;
#pragma unroll
    for (int j = 1; j < S; ++j) rtot += rr[j - 1];
#pragma unroll
    for (int k = 0; k < NQ; ++k) {
      float4 rs, rp;
      uint2 h, lw;
      round4<PREC>(sh[k], rs, h, lw);
      round4<PREC>(pr[k], rp, h, lw);
      // g = w_rank (r rho - y) + sum_j rr_j (r - r_j)
      float4 a = make_float4(fmaf(rs.x, rtot, -p.w_rank * ysh[k].x), fmaf(rs.y, rtot, -p.w_rank * ysh[k].y),
                             fmaf(rs.z, rtot, -p.w_rank * ysh[k].z), fmaf(rs.w, rtot, -p.w_rank * ysh[k].w));
      float4 c = make_float4(fmaf(rp.x, rtot, -p.w_rank * ypr[k].x), fmaf(rp.y, rtot, -p.w_rank * ypr[k].y),
                             fmaf(rp.z, rtot, -p.w_rank * ypr[k].z), fmaf(rp.w, rtot, -p.w_rank * ypr[k].w));
#pragma unroll
      for (int j = 1; j < S; ++j) {
        const float w = -rr[j - 1];
        a.x = fmaf(w, __shfl_xor_sync(0xffffffffu, rs.x, j * LPR), a.x);
        a.y = fmaf(w, __shfl_xor_sync(0xffffffffu, rs.y, j * LPR), a.y);
        a.z = fmaf(w, __shfl_xor_sync(0xffffffffu, rs.z, j * LPR), a.z);
        a.w = fmaf(w, __shfl_xor_sync(0xffffffffu, rs.w, j * LPR), a.w);
        c.x = fmaf(w, __shfl_xor_sync(0xffffffffu, rp.x, j * LPR), c.x);
        c.y = fmaf(w, __shfl_xor_sync(0xffffffffu, rp.y, j * LPR), c.y);
        c.z = fmaf(w, __shfl_xor_sync(0xffffffffu, rp.z, j * LPR), c.z);
        c.w = fmaf(w, __shfl_xor_sync(0xffffffffu, rp.w, j * LPR), c.w);
      }
      gsh[k] = a; gpr[k] = c;
    }
  }

  // ---- InfoNCE: operands 2t (shared half) and 2t + 1 (private half) of this tensor, one half at a time.  All loads of
  // a phase are issued together (the accumulator row, the positive-pair operand row; then the rows of the secondary
  // stream-K pieces): loads inside data-dependent loops would go out one round trip at a time.
  if (p.terms & FOCAL_TERM_NCE) {
    const uint64_t rowN = (uint64_t)g * p.bpad + I, rowsNce = (uint64_t)S * p.bpad;
    const float inv_tsn = 1.f / (p.T * (float)S * (float)(2 * p.b));
    const float inv_alpha = 1.f / p.alpha;
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      float4 tmp[NQ];
#pragma unroll
      for (int k = 0; k < NQ; ++k) tmp[k] = make_float4(0.f, 0.f, 0.f, 0.f);
      const OpDesc& op = p.ops[2 * t + half];
      const int wp = op.kb * p.epb;
      for (int u = 0; u < op.nuse; ++u) {
        const int qp = op.use_prob[u], side = op.use_side[u];
        const ProbDesc& prb = p.probs[qp];
        const OpDesc& po = p.ops[op.use_partner[u]];                  // partner operand: positive row p(k), same row index
        const uint8_t* pop = ws + po.off;
        const uint64_t lo_img = (uint64_t)(po.kb / 2) * rowsNce * 128;
        const uint64_t arow = ((uint64_t)side * S * p.bpad + rowN) * wp;
        const float* accp = reinterpret_cast<const float*>(ws + prb.dz_off) + arow;
        const uint64_t fidx = ((uint64_t)qp * S + g) * 2 + side;
        const int extra = __ldg(reinterpret_cast<const int32_t*>(ws + p.flag_nce_off) + fidx * (p.bpad / kTileM) + I / kTileM);
        const float* rs = reinterpret_cast<const float*>(ws + p.rsum_off) + ((uint64_t)(qp * S + g) * 2) * p.bpad;
        const float r_k = __ldg(rs + (uint64_t)side * p.bpad + I), r_p = __ldg(rs + (uint64_t)(1 - side) * p.bpad + I);
        const float gpos = __ldg(reinterpret_cast<const float*>(ws + p.pos_off) + fidx * p.bpad + I);
        float4 acc[NQ], zp[NQ];
#pragma unroll
        for (int k = 0; k < NQ; ++k) acc[k] = ldg4(accp + 4 * (k * LPR + l));
#pragma unroll
        for (int k = 0; k < NQ; ++k) zp[k] = ld_op4<PREC>(pop, lo_img, op_off8(rowsNce, rowN, 4 * (k * LPR + l)));
        for (int e = 1; e <= extra; ++e) {
          const float* acc2 = reinterpret_cast<const float*>(ws + prb.dz_off + e * p.dz2_delta) + arow;
          float4 a2[NQ];
#pragma unroll
          for (int k = 0; k < NQ; ++k) a2[k] = ldg4(acc2 + 4 * (k * LPR + l));
#pragma unroll
          for (int k = 0; k < NQ; ++k) { acc[k].x += a2[k].x; acc[k].y += a2[k].y; acc[k].z += a2[k].z; acc[k].w += a2[k].w; }
        }
        // positive column in fp32 (masked out of the tiles): W_kp - 2 is a tiny difference when the positive dominates;
        // its logit is the one the row-sum tile of this row saw
        const float wkp2 = ex2_approx(gpos) * (__frcp_rn(r_k) + __frcp_rn(r_p)) - 2.f;
        const float wq = prb.weight * inv_tsn * inv_alpha;
#pragma unroll
        for (int k = 0; k < NQ; ++k) {
          tmp[k].x = fmaf(wq, fmaf(wkp2, zp[k].x, acc[k].x), tmp[k].x);
          tmp[k].y = fmaf(wq, fmaf(wkp2, zp[k].y, acc[k].y), tmp[k].y);
          tmp[k].z = fmaf(wq, fmaf(wkp2, zp[k].z, acc[k].z), tmp[k].z);
          tmp[k].w = fmaf(wq, fmaf(wkp2, zp[k].w, acc[k].w), tmp[k].w);
        }
      }
      // d zh / d z = (I - zh zh^T) / n
      float dot[1] = {0.f};
#pragma unroll
      for (int k = 0; k < NQ; ++k) {
        const float4 xk = half ? pr[k] : sh[k];
        dot[0] = fmaf(tmp[k].x, xk.x, fmaf(tmp[k].y, xk.y, fmaf(tmp[k].z, xk.z, fmaf(tmp[k].w, xk.w, dot[0]))));
      }
      group_sum_n<LPR, 1>(dot);
      const float inrm = fminf(rsqrtf(half ? nb : na), 1.f / kNceEps);     // 1 / max(|z|, eps)
      const float dd = dot[0] * inrm * inrm;
#pragma unroll
      for (int k = 0; k < NQ; ++k) {
        const float4 xk = half ? pr[k] : sh[k];
        float4& gk = half ? gpr[k] : gsh[k];
        gk.x = fmaf(fmaf(-dd, xk.x, tmp[k].x), inrm, gk.x);
        gk.y = fmaf(fmaf(-dd, xk.y, tmp[k].y), inrm, gk.y);
        gk.z = fmaf(fmaf(-dd, xk.z, tmp[k].z), inrm, gk.z);
        gk.w = fmaf(fmaf(-dd, xk.w, tmp[k].w), inrm, gk.w);
      }
    }
  }

  // ---- orthogonality: (shared_t, private_t) and (private_t, private_t') for every other t' of the view
  if (orth_on) {
    const float a = p.w_orth / (float)p.B;
    {
      float dot[1] = {0.f};
#pragma unroll
      for (int k = 0; k < NQ; ++k)
        dot[0] = fmaf(sh[k].x, pr[k].x, fmaf(sh[k].y, pr[k].y, fmaf(sh[k].z, pr[k].z, fmaf(sh[k].w, pr[k].w, dot[0]))));
      group_sum_n<LPR, 1>(dot);
      const float nu = na + kOrthEps, nv = nb + kOrthEps;
      const float inv_den = rsqrtf(nu * nv), cs = dot[0] * inv_den;
      const float on = cs >= 0.f ? a : 0.f;                 // clamp_min passes gradient at equality
      const float ad = on * inv_den, au = -on * cs * __frcp_rn(nu), av = -on * cs * __frcp_rn(nv);
#pragma unroll
      for (int k = 0; k < NQ; ++k) {
        gsh[k].x = fmaf(ad, pr[k].x, fmaf(au, sh[k].x, gsh[k].x));
        gsh[k].y = fmaf(ad, pr[k].y, fmaf(au, sh[k].y, gsh[k].y));
        gsh[k].z = fmaf(ad, pr[k].z, fmaf(au, sh[k].z, gsh[k].z));
        gsh[k].w = fmaf(ad, pr[k].w, fmaf(au, sh[k].w, gsh[k].w));
        gpr[k].x = fmaf(ad, sh[k].x, fmaf(av, pr[k].x, gpr[k].x));
        gpr[k].y = fmaf(ad, sh[k].y, fmaf(av, pr[k].y, gpr[k].y));
        gpr[k].z = fmaf(ad, sh[k].z, fmaf(av, pr[k].z, gpr[k].z));
        gpr[k].w = fmaf(ad, sh[k].w, fmaf(av, pr[k].w, gpr[k].w));
      }
    }
    const int v0 = (t / p.M) * p.M;
    for (int t2 = v0; t2 < v0 + p.M; ++t2) {
      if (t2 == t) continue;
      const float* other = prv + ((size_t)(q * nT + t2) * S + g) * d;
      float4 v[NQ];
      float dot[1] = {0.f};
#pragma unroll
      for (int k = 0; k < NQ; ++k) {
        v[k] = ld4(other + 4 * (k * LPR + l));
        dot[0] = fmaf(pr[k].x, v[k].x, fmaf(pr[k].y, v[k].y, fmaf(pr[k].z, v[k].z, fmaf(pr[k].w, v[k].w, dot[0]))));
      }
      group_sum_n<LPR, 1>(dot);
      const float nb2 = __ldg(reinterpret_cast<const float*>(ws + p.nrm_off + ((uint64_t)t2 * p.Bpad + i) * 8) + 1);
      const float nu = nb + kOrthEps, nv = nb2 + kOrthEps;
      const float inv_den = rsqrtf(nu * nv), cs = dot[0] * inv_den;
      const float on = cs >= 0.f ? a : 0.f;
      const float ad = on * inv_den, au = -on * cs * __frcp_rn(nu);
#pragma unroll
      for (int k = 0; k < NQ; ++k) {
        gpr[k].x = fmaf(ad, v[k].x, fmaf(au, pr[k].x, gpr[k].x));
        gpr[k].y = fmaf(ad, v[k].y, fmaf(au, pr[k].y, gpr[k].y));
        gpr[k].z = fmaf(ad, v[k].z, fmaf(au, pr[k].z, gpr[k].z));
        gpr[k].w = fmaf(ad, v[k].w, fmaf(au, pr[k].w, gpr[k].w));
      }
    }
  }
  float* out = grad_base(p, gp, ws, t) + (size_t)i * p.D;
#pragma unroll
  for (int k = 0; k < NQ; ++k) {
    if (FB_L2_HINTS & 2) st4_hint(out + 4 * (k * LPR + l), gsh[k], l2_policy_evict_first());
    else *reinterpret_cast<float4*>(out + 4 * (k * LPR + l)) = gsh[k];
  }
#pragma unroll
  for (int k = 0; k < NQ; ++k) {
    if (FB_L2_HINTS & 2) st4_hint(out + d + 4 * (k * LPR + l), gpr[k], l2_policy_evict_first());
    else *reinterpret_cast<float4*>(out + d + 4 * (k * LPR + l)) = gpr[k];
  }
}

}  // namespace fb
