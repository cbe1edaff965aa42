// Single-pass row kernels (HBM-bound): prologue, intra-sequence distances, log-row-sums, gradient assembly,
// loss reduction.  One warp per row, rows staged in shared memory, warp-shuffle reductions, coalesced stores.
#pragma once
#include "peer.cuh"
#include "plan.h"
#include "ptx.cuh"

namespace fb {

struct FeatPtrs {
  const float* x[kMaxT];
};
struct GradPtrs {
  float* g[kMaxT];
};

// caller pointers: kernel arguments, or (indirect mode) the table focal_b200_set_ptrs wrote into the workspace
__device__ __forceinline__ const float* feat_base(const Plan& p, const FeatPtrs& f, const uint8_t* ws, int t) {
  if (!p.indirect) return f.x[t];
  return reinterpret_cast<const float*>(
      __ldg(reinterpret_cast<const unsigned long long*>(&reinterpret_cast<const PtrTable*>(ws + p.ptrs_off)->x[t])));
}
__device__ __forceinline__ float* grad_base(const Plan& p, const GradPtrs& g, const uint8_t* ws, int t) {
  if (!p.indirect) return g.g[t];
  return reinterpret_cast<float*>(
      __ldg(reinterpret_cast<const unsigned long long*>(&reinterpret_cast<const PtrTable*>(ws + p.ptrs_off)->g[t])));
}
__global__ void set_ptrs_kernel(PtrTable* __restrict__ dst, const PtrTable src) {
  if (threadIdx.x == 0) *dst = src;
}

constexpr float kNceEps = 1e-8f;    // nn.CosineSimilarity eps (loss.py:15)
constexpr float kOrthEps = 1e-12f;  // cosine_embedding_loss EPSILON (loss.py:16)

// Cooperative (one warp) load of the nT rows `i` into shared memory: xs[t * D + c].
__device__ __forceinline__ void load_rows(const Plan& p, const FeatPtrs& f, const uint8_t* ws, int i, float* xs,
                                          int lane) {
  const int D = p.D;
  if ((D & 3) == 0) {
    const int nv = D >> 2;
    for (int t = 0; t < p.nT; ++t) {
      const float4* src = reinterpret_cast<const float4*>(feat_base(p, f, ws, t) + feat_row_off(p, i));
      float4* dst = reinterpret_cast<float4*>(xs + t * D);
      for (int c = lane; c < nv; c += 32) dst[c] = __ldg(src + c);
    }
  } else {
    for (int t = 0; t < p.nT; ++t)
      for (int c = lane; c < D; c += 32) xs[t * D + c] = __ldg(feat_base(p, f, ws, t) + feat_row_off(p, i) + c);
  }
  __syncwarp();
}

// fp32 -> bf16 (round to nearest even) -> fp32: the value the tensor core sees for an operand element
__device__ __forceinline__ float bf16_round(float v) {
  return __uint_as_float(pack_bf16x2(v, 0.f) << 16);
}

// split tiles (fp32 mode): hi = bf16(v), lo = bf16(v - hi); the tensor core sees hi + lo (exact in fp32: 16 bits)
__device__ __forceinline__ float split_round(float v) {
  const float hi = bf16_round(v);
  return hi + bf16_round(v - hi);
}
// the value the tensor core sees for an operand element in the given tile precision (FOCAL_PREC_*)
__device__ __forceinline__ float op_round(int prec, float v) { return prec ? split_round(v) : bf16_round(v); }
template <int PREC>
__device__ __forceinline__ float op_round_t(float v) { return PREC ? split_round(v) : bf16_round(v); }
// Squared operand element as the temporal distances need it: d^2 = n_i + n_j - 2 G_ij with G from the tiles.  bf16 tiles:
// hi * hi, exactly the tile's own product, so coincident rows give d^2 = 0.  Split tiles: the full (hi + lo)^2.  The tiles
// never form lo * lo; in G that term is zero-mean noise (2^-18 relative), but in a squared norm it is a sum of squares:
// leaving it out made every distance 3.1e-5 too small (measured; profiles/r2_accuracy.txt), a systematic bias of the
// hinge term.  The price is d^2 = 2 sum(lo^2) ~ 1e-3 instead of 0 for coincident rows -- the size of the reference's own
// fp32 cancellation error in torch.cdist's matmul form.
__device__ __forceinline__ float tile_sq(int prec, float a) {
  const float ha = bf16_round(a);
  if (!prec) return ha * ha;
  const float r = ha + bf16_round(a - ha);
  return r * r;
}

__device__ __forceinline__ float warp_dot(const float* a, const float* b, int n, int lane) {
  float s = 0.f;
  for (int c = lane; c < n; c += 32) s = fmaf(a[c], b[c], s);
  return warp_sum(s);
}

// ---------------------------------------------------------------------------------------------------------
// K0: zero the padding rows of the operand tiles and per-row vectors (only launched when padding exists)
// ---------------------------------------------------------------------------------------------------------
// `tid` of `stride` threads; only the padding is visited (an earlier version scanned every row of the operand arrays and
// cost 18 us per step at the headline shape, where the padding is 128 rows).
__device__ __forceinline__ void zero_pad_body(const Plan& p, uint8_t* __restrict__ ws, long tid, long stride) {
  const int padN = p.bpad - p.b;      // per (op, kb, s)
  const int padT = p.Bpad - p.Bt;     // per (tensor, kb)
  const uint4 z = make_uint4(0, 0, 0, 0);
  if (padN > 0) {
    for (int o = 0; o < p.nOps; ++o) {
      const long chunks = (long)p.ops[o].kb * p.S * padN * 8;
      for (long e = tid; e < chunks; e += stride) {
        const int c = e & 7;
        long r = e >> 3;
        const int pr = r % padN; r /= padN;
        const int s = r % p.S; r /= p.S;
        const int kb = (int)r;
        const uint64_t row = (uint64_t)s * p.bpad + p.b + pr;
        *reinterpret_cast<uint4*>(ws + p.ops[o].off + ((uint64_t)kb * p.S * p.bpad + row) * 128 + c * 16) = z;
      }
    }
    float* rinv = reinterpret_cast<float*>(ws + p.rinv_off);
    float* rsum = reinterpret_cast<float*>(ws + p.rsum_off);
    const long nr = (long)p.nProb * p.S * 2 * padN;
    for (long e = tid; e < nr; e += stride) {
      const long grp = e / padN;
      const int pr = e % padN;
      rinv[grp * p.bpad + p.b + pr] = 0.f;
      rsum[grp * p.bpad + p.b + pr] = 1.f;
    }
  }
  float* sq = reinterpret_cast<float*>(ws + p.sq_off);
  float* mi = reinterpret_cast<float*>(ws + p.mintra_off);
  if (p.Sp != p.S) {
    // padded sequences (seq_len not a power of two): phantom rows (position >= S) are spread over the whole temporal row
    // space, so every row is looked at
    const long chunks = (long)p.nT * p.kbFull * p.Bpad * 8;
    for (long e = tid; e < chunks; e += stride) {
      const int c = e & 7;
      long r = e >> 3;
      const int row = r % p.Bpad; r /= p.Bpad;
      if (row < p.Bt && (row % p.Sp) < p.S) continue;
      const uint64_t tk = r;      // t * kbFull + kb
      *reinterpret_cast<uint4*>(ws + p.xt_off + (tk * p.Bpad + row) * 128 + c * 16) = z;
    }
    for (long e = tid; e < (long)p.nT * p.Bpad; e += stride) {
      const int row = e % p.Bpad;
      if (row < p.Bt && (row % p.Sp) < p.S) continue;
      sq[e] = 0.f;
      mi[e] = 0.f;
    }
  } else if (padT > 0) {
    // temporal row space: the rows beyond the batch, [Bt, Bpad) of every (tensor, K block)
    const long chunks = (long)p.nT * p.kbFull * padT * 8;
    for (long e = tid; e < chunks; e += stride) {
      const int c = e & 7;
      long r = e >> 3;
      const int row = p.Bt + (int)(r % padT); r /= padT;
      const uint64_t tk = r;      // t * kbFull + kb
      *reinterpret_cast<uint4*>(ws + p.xt_off + (tk * p.Bpad + row) * 128 + c * 16) = z;
    }
    for (long e = tid; e < (long)p.nT * padT; e += stride) {
      const long idx = (e / padT) * p.Bpad + p.Bt + (e % padT);
      sq[idx] = 0.f;
      mi[idx] = 0.f;
    }
  }
}
__global__ void zero_pad_kernel(const __grid_constant__ Plan p, uint8_t* __restrict__ ws) {
  zero_pad_body(p, ws, (long)blockIdx.x * blockDim.x + threadIdx.x, (long)gridDim.x * blockDim.x);
}

// ---------------------------------------------------------------------------------------------------------
// K1: prologue.  One warp per row i of the batch, all 2M tensors of that row.
//   - normalised, pre-scaled bf16 InfoNCE operands (position-major rows, swizzled 128-B K blocks)
//   - bf16 temporal operands (natural rows) + squared norms of the ROUNDED values
//   - owned rows: positive-pair logits (loss.py:75-79) and orthogonality terms (loss.py:89-106)
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(32 * kRowsPerBlock) prologue_kernel(const __grid_constant__ Plan p,
                                                                      const __grid_constant__ FeatPtrs f,
                                                                      uint8_t* __restrict__ ws) {
  extern __shared__ float smem_f[];
  __shared__ float red[kRowsPerBlock][4];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int i = blockIdx.x * kRowsPerBlock + warp;
  float* xs = smem_f + (size_t)warp * p.nT * p.D;
  float acc_orth = 0.f, acc_ps = 0.f, acc_pp = 0.f;
  if (i < p.B) {
    load_rows(p, f, ws, i, xs, lane);
    const int D = p.D, d = p.d;
    const int I = i / p.S, s = i % p.S;
    const uint64_t rowN = (uint64_t)s * p.bpad + I;          // position-major row of the InfoNCE operands
    const uint64_t rowsNce = (uint64_t)p.S * p.bpad;

    // squared norms of the shared half, the private half and the odd leftover column
    float ssh[kMaxT], spr[kMaxT], sfull[kMaxT];
    for (int t = 0; t < p.nT; ++t) {
      const float* x = xs + t * D;
      float a = 0.f, b = 0.f, c = 0.f;
      for (int k = lane; k < d; k += 32) { a = fmaf(x[k], x[k], a); b = fmaf(x[d + k], x[d + k], b); }
      if ((D & 1) && lane == 0) c = x[D - 1] * x[D - 1];
      a = warp_sum(a); b = warp_sum(b); c = warp_sum(c);
      ssh[t] = a; spr[t] = b; sfull[t] = a + b + c;
    }

    // ---- InfoNCE operands
    if (p.terms & FOCAL_TERM_NCE) {
      for (int o = 0; o < p.nOps; ++o) {
        const OpDesc& op = p.ops[o];
        const float ss = (op.width == D) ? sfull[op.tensor] : (op.col0 == 0 ? ssh[op.tensor] : spr[op.tensor]);
        const float scale = p.alpha / fmaxf(sqrtf(ss), kNceEps);
        const float* x = xs + op.tensor * D + op.col0;
        const int kh = p.prec == FOCAL_PREC_FP32 ? op.kb / 2 : op.kb;       // blocks of one image
        for (int kb = 0; kb < kh; ++kb) {                                      // 64 elements per K block: a pair per lane
          uint8_t* dst = ws + op.off + ((uint64_t)kb * rowsNce + rowN) * 128 + tile_byte_bf16((uint32_t)rowN, 2 * lane);
          const int e = kb * 64 + 2 * lane;
          const float v0 = (e < op.width) ? x[e] * scale : 0.f;
          const float v1 = (e + 1 < op.width) ? x[e + 1] * scale : 0.f;
          const uint32_t hi = pack_bf16x2(v0, v1);
          *reinterpret_cast<uint32_t*>(dst) = hi;
          if (p.prec == FOCAL_PREC_FP32)                                       // lo image: kh blocks further
            *reinterpret_cast<uint32_t*>(dst + (uint64_t)kh * rowsNce * 128) =
                pack_bf16x2(v0 - __uint_as_float(hi << 16), v1 - __uint_as_float(hi & 0xffff0000u));
        }
      }
    }
    // ---- temporal operands (row tr(i) of the temporal row space)
    if (p.terms & FOCAL_TERM_TEMPORAL) {
      const int it = tmp_row(p, i);
      for (int t = 0; t < p.nT; ++t) {
        const float* x = xs + t * D;
        float sq = 0.f;
        const int kh = p.prec == FOCAL_PREC_FP32 ? p.kbFull / 2 : p.kbFull;
        for (int kb = 0; kb < kh; ++kb) {
          uint8_t* dst = ws + p.xt_off + (((uint64_t)t * p.kbFull + kb) * p.Bpad + it) * 128 + tile_byte_bf16((uint32_t)it, 2 * lane);
          const int e = kb * 64 + 2 * lane;
          const float v0 = (e < D) ? x[e] : 0.f;
          const float v1 = (e + 1 < D) ? x[e + 1] : 0.f;
          const uint32_t pk = pack_bf16x2(v0, v1);
          const float r0 = __uint_as_float(pk << 16), r1 = __uint_as_float(pk & 0xffff0000u);
          *reinterpret_cast<uint32_t*>(dst) = pk;
          float l0 = 0.f, l1 = 0.f;
          if (p.prec == FOCAL_PREC_FP32) {
            const uint32_t lo = pack_bf16x2(v0 - r0, v1 - r1);
            *reinterpret_cast<uint32_t*>(dst + (uint64_t)kh * p.Bpad * 128) = lo;
            l0 = __uint_as_float(lo << 16); l1 = __uint_as_float(lo & 0xffff0000u);
          }
          sq = fmaf(r0 + l0, r0 + l0, fmaf(r1 + l1, r1 + l1, sq));        // = tile_sq(x)
        }
        sq = warp_sum(sq);
        if (lane == 0) reinterpret_cast<float*>(ws + p.sq_off)[(uint64_t)t * p.Bpad + it] = sq;
      }
    }
    // ---- loss terms of the owned rows
    if (I >= p.seq0 && I < p.seq1) {
      if (p.terms & FOCAL_TERM_ORTH) {
        for (int k = 0; k < p.nOrth; ++k) {
          const OrthDesc& od = p.orth[k];
          const float dot = warp_dot(xs + od.tu * D + od.cu, xs + od.tv * D + od.cv, od.width, lane);
          const float nu = (od.cu == 0 ? ssh[od.tu] : spr[od.tu]) + kOrthEps;
          const float nv = (od.cv == 0 ? ssh[od.tv] : spr[od.tv]) + kOrthEps;
          const float cs = dot / sqrtf(nu * nv);
          acc_orth += fmaxf(cs, 0.f);
        }
      }
    }
  }
  if (lane == 0) { red[warp][0] = acc_orth; red[warp][1] = acc_ps; red[warp][2] = acc_pp; }
  __syncthreads();
  if (threadIdx.x < 3) {
    float s = 0.f;
    for (int w = 0; w < kRowsPerBlock; ++w) s += red[w][threadIdx.x];
    if (threadIdx.x == 0) s /= (float)p.B;
    reinterpret_cast<float*>(ws + p.part1_off)[(size_t)blockIdx.x * 4 + threadIdx.x] = s;
  }
}

// ---------------------------------------------------------------------------------------------------------
// K1b: intra-sequence mean distance m_II (loss.py:118-124, block diagonal of the block means), direct
// differences in fp32 of the SAME bf16-rounded rows the distance tiles use: the whole temporal term is then the
// exact loss of the rounded features, and rounding noise that is coherent per row cancels between m_II and m_IJ.
// One warp per (tensor, sequence); written per row so the Gram kernel can TMA it as a column vector.
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) intra_kernel(const __grid_constant__ Plan p, const __grid_constant__ FeatPtrs f,
                                                    uint8_t* __restrict__ ws) {
  const int lane = threadIdx.x & 31;
  const long w = (long)blockIdx.x * 4 + (threadIdx.x >> 5);
  if (w >= (long)p.nT * p.b) return;
  const int t = (int)(w / p.b), I = (int)(w % p.b);
  const int S = p.S, D = p.D;
  const float* x = feat_base(p, f, ws, t) + feat_row_off(p, I * S);
  float tot = 0.f;
  for (int a = 0; a < S; ++a)
    for (int b2 = a + 1; b2 < S; ++b2) {
      float d2 = 0.f;
      for (int c = lane; c < D; c += 32) {
        const float df = op_round(p.prec, __ldg(x + a * D + c)) - op_round(p.prec, __ldg(x + b2 * D + c));
        d2 = fmaf(df, df, d2);
      }
      d2 = warp_sum(d2);
      tot += sqrtf(d2);
    }
  const float m = 2.f * tot / (float)(S * S - S);
  float* mi = reinterpret_cast<float*>(ws + p.mintra_off) + (uint64_t)t * p.Bpad + (uint64_t)I * p.Sp;
  for (int a = lane; a < S; a += 32) mi[a] = m;
}

// ---------------------------------------------------------------------------------------------------------
// K2r: reduce the column-split row sums of the owned rows, publish r, 1/r and the log-sum partials.
// all_rows != 0: only rebuild 1/r for every valid row from r (after the multi-GPU exchange of r).
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) nce_lse_kernel(const __grid_constant__ Plan p,
                                                      const __grid_constant__ PeerWs pw, uint8_t* __restrict__ ws,
                                                      int all_rows) {
  pdl_launch_dependents();
  pdl_wait();
  __shared__ float red[8][2];
  const long n = (long)p.nProb * p.S * 2 * p.bpad;
  float ls = 0.f, lp = 0.f;
  if (all_rows) {
    const long e = (long)blockIdx.x * 256 + threadIdx.x;
    if (e < n && (int)(e % p.bpad) < p.b)
      reinterpret_cast<float*>(ws + p.rinv_off)[e] = 1.f / reinterpret_cast<const float*>(ws + p.rsum_off)[e];
  } else {
    // the grid covers the owned sequences only: thread -> ((problem, position, side), k in [seq0, seq1))
    const int per = p.seq1 - p.seq0;
    const long t = (long)blockIdx.x * 256 + threadIdx.x;
    if (t < (long)p.nProb * p.S * 2 * per) {
      const long grp = t / per;
      const long e = grp * p.bpad + p.seq0 + (int)(t - grp * per);
      const int q = (int)(grp / (p.S * 2));
      const float* rp = reinterpret_cast<const float*>(ws + p.rpart_off);
      float r = 0.f;
      for (int sp = 0; sp < p.nsplit_fwd; ++sp) r += rp[(long)sp * n + e];
      // row-sharded jobs: every rank needs r and 1/r of every row -> store into all workspaces (pw.world == 1 otherwise)
      if (pw.mc) {                      // NVSwitch multicast: one store serves every rank
        mc_st4(pw.mc + p.rsum_off + (uint64_t)e * 4, r);
        mc_st4(pw.mc + p.rinv_off + (uint64_t)e * 4, 1.f / r);
      } else for (int rk = 0; rk < pw.world; ++rk) {
        reinterpret_cast<float*>(pw.ws[rk] + p.rsum_off)[e] = r;
        reinterpret_cast<float*>(pw.ws[rk] + p.rinv_off)[e] = 1.f / r;
      }
      // row loss ln sum_{j != k} exp(s_kj) - s_{k,p(k)}, both from the tiles of THIS row (log2-domain logit G: s = ln2 G)
      const float g_pos = reinterpret_cast<const float*>(ws + p.pos_off)[e];
      const float lg = (logf(r) - 0.6931471805599453f * g_pos) / ((float)p.S * (float)(2 * p.b));
      if (p.probs[q].kind == 0) ls = lg; else lp = lg;
    }
  }
  if (!all_rows) {
    ls = warp_sum(ls); lp = warp_sum(lp);
    if ((threadIdx.x & 31) == 0) { red[threadIdx.x >> 5][0] = ls; red[threadIdx.x >> 5][1] = lp; }
    __syncthreads();
    if (threadIdx.x < 2) {
      float s = 0.f;
      for (int w = 0; w < 8; ++w) s += red[w][threadIdx.x];
      reinterpret_cast<float*>(ws + p.part2_off)[(size_t)blockIdx.x * 2 + threadIdx.x] = s;
    }
    if (pw.world > 1) peer_epoch_bump(p, ws);      // row sums of the owned rows are out
  }
}

// ---------------------------------------------------------------------------------------------------------
// K4: gradient assembly.  One warp per owned row i; everything that is O(B D) happens here in fp32:
//   InfoNCE: d/dz of the normalisation (I - zh zh^T)/n applied to the accumulated W Z_J, positive-pair term
//   temporal: x_i rho_i - (R X)_i  +  exact intra-sequence part
//   orthogonality: closed form
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(32 * kRowsPerBlock) finalize_kernel(const __grid_constant__ Plan p,
                                                                      const __grid_constant__ FeatPtrs f,
                                                                      const __grid_constant__ GradPtrs g,
                                                                      const uint8_t* __restrict__ ws) {
  extern __shared__ float smem_f[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int i = p.seq0 * p.S + blockIdx.x * kRowsPerBlock + warp;
  if (i >= p.seq1 * p.S) return;
  const int D = p.D, d = p.d, S = p.S;
  float* xs = smem_f + (size_t)warp * (2 * p.nT + 1) * D;
  float* gs = xs + (size_t)p.nT * D;
  float* tmp = gs + (size_t)p.nT * D;
  load_rows(p, f, ws, i, xs, lane);
  for (int c = lane; c < p.nT * D; c += 32) gs[c] = 0.f;
  const int I = i / S, s = i % S;
  const uint64_t rowN = (uint64_t)s * p.bpad + I;

  float ssh[kMaxT], spr[kMaxT], sfull[kMaxT];
  for (int t = 0; t < p.nT; ++t) {
    const float* x = xs + t * D;
    float a = 0.f, b = 0.f, c = 0.f;
    for (int k = lane; k < d; k += 32) { a = fmaf(x[k], x[k], a); b = fmaf(x[d + k], x[d + k], b); }
    if ((D & 1) && lane == 0) c = x[D - 1] * x[D - 1];
    a = warp_sum(a); b = warp_sum(b); c = warp_sum(c);
    ssh[t] = a; spr[t] = b; sfull[t] = a + b + c;
  }
  __syncwarp();

  // ---- temporal
  if ((p.terms & FOCAL_TERM_TEMPORAL) && p.b > 1 && S > 1) {
    const float bb = (float)p.b * (float)(p.b - 1);
    const int Dp = p.kbFull * p.epb;
    for (int t = 0; t < p.nT; ++t) {
      const float* x = xs + t * D;
      // a row block whose column tiles were split over several CTAs (stream-K) has one set of accumulators per piece
      const int it = tmp_row(p, i);
      const int extra = reinterpret_cast<const int32_t*>(ws + p.flag_tmp_off)[(uint64_t)t * (p.Bpad / kTileM) + it / kTileM];
      float rho = reinterpret_cast<const float*>(ws + p.rho_off)[(uint64_t)t * p.Bpad + it];
      const float* y = reinterpret_cast<const float*>(ws + p.dx_off) + ((uint64_t)t * p.Bpad + it) * Dp;
      int cnt = reinterpret_cast<const int32_t*>(ws + p.cnt_off)[(uint64_t)t * p.bpad + I];
      for (int k = 1; k <= extra; ++k) {
        rho += reinterpret_cast<const float*>(ws + p.rho_off + k * p.rho2_delta)[(uint64_t)t * p.Bpad + it];
        cnt += reinterpret_cast<const int32_t*>(ws + p.cnt_off + k * p.cnt2_delta)[(uint64_t)t * p.bpad + I];
      }
      for (int c = lane; c < D; c += 32) {
        float yc = y[c];
        for (int k = 1; k <= extra; ++k)
          yc += (reinterpret_cast<const float*>(ws + p.dx_off + k * p.dx2_delta) + ((uint64_t)t * p.Bpad + it) * Dp)[c];
        gs[t * D + c] = p.w_rank * (op_round(p.prec, x[c]) * rho - yc);
      }
      // intra-sequence pairs: dL/dm_II = cnt / (b(b-1)), spread over S^2 - S ordered pairs, both orders
      const float coef = p.w_rank * 2.f * (float)cnt / (bb * (float)(S * S - S));
      if (cnt > 0) {
        const float* base = feat_base(p, f, ws, t) + feat_row_off(p, I * S);
        for (int j = 0; j < S; ++j) {
          if (j == s) continue;
          float d2 = 0.f;
          for (int c = lane; c < D; c += 32) {
            const float df = op_round(p.prec, x[c]) - op_round(p.prec, __ldg(base + (size_t)j * D + c));
            d2 = fmaf(df, df, d2);
          }
          d2 = warp_sum(d2);
          if (d2 > 0.f) {
            const float r = coef * rsqrtf(d2);
            for (int c = lane; c < D; c += 32)
              gs[t * D + c] += r * (op_round(p.prec, x[c]) - op_round(p.prec, __ldg(base + (size_t)j * D + c)));
          }
        }
      }
    }
  }
  __syncwarp();

  // ---- InfoNCE
  if (p.terms & FOCAL_TERM_NCE) {
    const float inv_tsn = 1.f / (p.T * (float)S * (float)(2 * p.b));
    const float inv_alpha = 1.f / p.alpha;
    for (int o = 0; o < p.nOps; ++o) {
      const OpDesc& op = p.ops[o];
      const int w = op.width, wp = op.kb * p.epb;
      const float ss = (w == D) ? sfull[op.tensor] : (op.col0 == 0 ? ssh[op.tensor] : spr[op.tensor]);
      const float nrm = fmaxf(sqrtf(ss), kNceEps);
      const float* x = xs + op.tensor * D + op.col0;
      bool used = false;
      for (int c = lane; c < w; c += 32) tmp[c] = 0.f;
      for (int q = 0; q < p.nProb; ++q) {
        const ProbDesc& pr = p.probs[q];
        int side = -1;
        if (pr.opA == o) side = 0; else if (pr.opB == o) side = 1;
        if (side < 0) continue;
        used = true;
        const OpDesc& po = p.ops[side == 0 ? pr.opB : pr.opA];            // partner operand: positive row p(k)
        const float pss = (po.width == D) ? sfull[po.tensor] : (po.col0 == 0 ? ssh[po.tensor] : spr[po.tensor]);
        const float pinv = p.alpha / fmaxf(sqrtf(pss), kNceEps);
        const float* px = xs + po.tensor * D + po.col0;
        const float* acc = reinterpret_cast<const float*>(ws + pr.dz_off) +
                           ((uint64_t)side * S * p.bpad + rowN) * wp;
        const int extra = reinterpret_cast<const int32_t*>(ws + p.flag_nce_off)[(((uint64_t)q * S + s) * 2 + side) *
                                                                                (p.bpad / kTileM) + I / kTileM];
        const float wq = pr.weight * inv_tsn;
        // The positive column p(k) is masked out of the tiles and handled here in fp32: where the positive
        // dominates the row (small T, aligned views) W_kp - 2 is a tiny difference that bf16 W would destroy.
        const float gpos = reinterpret_cast<const float*>(ws + p.pos_off)[((uint64_t)(q * S + s) * 2 + side) * p.bpad + I];
        const float* rs = reinterpret_cast<const float*>(ws + p.rsum_off) + ((uint64_t)(q * S + s) * 2) * p.bpad;
        const float wkp = exp2f(gpos) * (1.f / rs[(uint64_t)side * p.bpad + I] + 1.f / rs[(uint64_t)(1 - side) * p.bpad + I]);
        for (int c = lane; c < w; c += 32) {
          float ac = acc[c];
          for (int k = 1; k <= extra; ++k)
            ac += (reinterpret_cast<const float*>(ws + pr.dz_off + k * p.dz2_delta) + ((uint64_t)side * S * p.bpad + rowN) * wp)[c];
          tmp[c] += wq * inv_alpha * (ac + (wkp - 2.f) * op_round(p.prec, px[c] * pinv));
        }
      }
      if (!used) continue;
      // d zh / d z = (I - zh zh^T) / n
      float dot = 0.f;
      for (int c = lane; c < w; c += 32) dot = fmaf(tmp[c], x[c], dot);
      dot = warp_sum(dot) / (nrm * nrm);
      float* go = gs + op.tensor * D + op.col0;
      for (int c = lane; c < w; c += 32) go[c] += (tmp[c] - dot * x[c]) / nrm;
      __syncwarp();
    }
  }

  // ---- orthogonality
  if (p.terms & FOCAL_TERM_ORTH) {
    for (int k = 0; k < p.nOrth; ++k) {
      const OrthDesc& od = p.orth[k];
      const float* u = xs + od.tu * D + od.cu;
      const float* v = xs + od.tv * D + od.cv;
      const float dot = warp_dot(u, v, od.width, lane);
      const float nu = (od.cu == 0 ? ssh[od.tu] : spr[od.tu]) + kOrthEps;
      const float nv = (od.cv == 0 ? ssh[od.tv] : spr[od.tv]) + kOrthEps;
      const float den = sqrtf(nu * nv);
      const float cs = dot / den;
      if (cs >= 0.f) {                                    // clamp_min passes gradient at equality
        const float a = p.w_orth / (float)p.B;
        float* gu = gs + od.tu * D + od.cu;
        float* gv = gs + od.tv * D + od.cv;
        for (int c = lane; c < od.width; c += 32) {
          const float uc = u[c], vc = v[c];
          gu[c] += a * (vc / den - cs * uc / nu);
          gv[c] += a * (uc / den - cs * vc / nv);
        }
      }
      __syncwarp();
    }
  }
  __syncwarp();
  for (int t = 0; t < p.nT; ++t) {
    float* out = grad_base(p, g, ws, t) + (size_t)i * D;
    for (int c = lane; c < D; c += 32) out[c] = gs[t * D + c];
  }
}

// ---------------------------------------------------------------------------------------------------------
// K5: deterministic reduction of the per-block partials -> {total, shared, private, orth, temporal}
// ---------------------------------------------------------------------------------------------------------
// Announce and / or wait phase alone, as a one-block launch: used instead of the ones at the head of the persistent Gram
// kernels when several ranks share one device (tests), where a spinning 148-CTA launch would keep the other ranks'
// kernels off the SMs.
__global__ void __launch_bounds__(32) peer_wait_kernel(const __grid_constant__ Plan p,
                                                       const __grid_constant__ PeerWs pw, int announce, int wait) {
  if (announce) peer_announce_epoch(p, pw.ws, pw.world, pw.rank);
  if (wait) peer_wait(pw.ws[pw.rank], p, pw.world, pw.rank);
}

// Sum of the loss partials of every kernel of the step, in double and in a fixed order (bit-reproducible); ranks of a
// row-sharded job exchange their four sums through peer memory.  Runs in ONE block of any size (all of its threads call
// it): its own launch (loss_reduce_kernel), or an extra block of finalize_v3_kernel.
__device__ __forceinline__ void loss_reduce_body(const Plan& p, const PeerWs& pw, uint8_t* __restrict__ ws,
                                                 float* __restrict__ loss5, int nce_blocks_valid, int temporal_nan) {
  __shared__ double red[256];
  const int tid = threadIdx.x, nthr = blockDim.x < 256 ? (int)blockDim.x : 256;
  const float* p1 = reinterpret_cast<const float*>(ws + p.part1_off);
  const float* p2 = reinterpret_cast<const float*>(ws + p.part2_off);
  const float* p3 = reinterpret_cast<const float*>(ws + p.part3_off);
  double acc[4] = {0, 0, 0, 0};     // shared, private, orth, temporal
  if (tid < nthr) {
    for (int k = tid; k < p.nblk1; k += nthr) {
      acc[2] += p1[(size_t)k * 4 + 0];
      acc[0] += p1[(size_t)k * 4 + 1];
      acc[1] += p1[(size_t)k * 4 + 2];
    }
    if (p.terms & FOCAL_TERM_NCE)
      for (int k = tid; k < nce_blocks_valid; k += nthr) {
        acc[0] += p2[(size_t)k * 2 + 0];
        acc[1] += p2[(size_t)k * 2 + 1];
      }
    if ((p.terms & FOCAL_TERM_TEMPORAL) && !temporal_nan) {
      const int t0 = (p.seq0 * p.Sp) / kTileM;
      const int nrt = (p.seq1 * p.Sp + kTileM - 1) / kTileM - t0;
      for (int k = tid; k < p.np_tmp * p.nT * nrt; k += nthr) acc[3] += p3[k];
    }
  }
  double out[4];
  for (int a = 0; a < 4; ++a) {
    for (int k = tid; k < 256; k += blockDim.x) red[k] = 0.0;
    __syncthreads();
    if (tid < nthr) red[tid] = acc[a];
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
      if (tid < o) red[tid] += red[tid + o];
      __syncthreads();
    }
    out[a] = red[0];
    __syncthreads();
  }
  if (pw.world > 1) {
    // all-reduce over the ranks: publish the partial sums of the owned rows, barrier, add them in rank order
    if (tid < pw.world) {
      double* slot = reinterpret_cast<double*>(pw.ws[tid] + p.lossx_off) + pw.rank * 8;
      for (int a = 0; a < 4; ++a) slot[a] = out[a];
    }
    peer_barrier(p, pw);
    const volatile double* all = reinterpret_cast<const volatile double*>(ws + p.lossx_off);
    for (int a = 0; a < 4; ++a) {
      double s = 0;
      for (int r = 0; r < pw.world; ++r) s += all[r * 8 + a];
      out[a] = s;
    }
  }
  if (tid == 0) {
    if (p.indirect) loss5 = reinterpret_cast<const PtrTable*>(ws + p.ptrs_off)->loss5;
    if (temporal_nan && (p.terms & FOCAL_TERM_TEMPORAL)) out[3] = __longlong_as_double(0x7ff8000000000000LL);
    const double total = (double)p.w_shared * out[0] + (double)p.w_private * out[1] + (double)p.w_orth * out[2] +
                         (double)p.w_rank * out[3];
    loss5[0] = (float)total; loss5[1] = (float)out[0]; loss5[2] = (float)out[1];
    loss5[3] = (float)out[2]; loss5[4] = (float)out[3];
    double* ld = reinterpret_cast<double*>(ws + p.lossd_off);
    ld[0] = total; ld[1] = out[0]; ld[2] = out[1]; ld[3] = out[2]; ld[4] = out[3];
  }
}
__global__ void __launch_bounds__(256) loss_reduce_kernel(const __grid_constant__ Plan p,
                                                          const __grid_constant__ PeerWs pw, uint8_t* __restrict__ ws,
                                                          float* __restrict__ loss5, int nce_blocks_valid,
                                                          int temporal_nan) {
  loss_reduce_body(p, pw, ws, loss5, nce_blocks_valid, temporal_nan);
}

}  // namespace fb
