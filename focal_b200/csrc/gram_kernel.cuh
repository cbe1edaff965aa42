// Flash-style Gram kernels on tcgen05 / TMEM (sm_100a).
//
// One persistent, warp-specialised kernel template serves the four dense passes of the FOCAL loss:
//
//   NCE_FWD   G = Z_I Z_J^T (log2-domain logits)  -> row sums  sum_{j != k} 2^G            (loss.py:74-85)
//   NCE_BWD   recompute G, W = 2^G (1/r_k + 1/r_j) -> O += W Z_J  (second UMMA, W from TMEM) (autograd of the above)
//   TMP_FWD   G = X_I X_J^T -> delta = sqrt(n_i+n_j-2G) -> S x S block means -> hinge        (loss.py:113-135)
//   TMP_BWD   same pass + r_ij = coef_IJ / delta_ij -> O += R X_J                           (fwd + bwd fused)
//
// The N x N logit / distance matrices are never written to memory.  Operands are bf16 tiles that the prologue
// laid out in HBM exactly as the UMMA wants them in shared memory (K-block-major, 128-byte rows, SWIZZLE_128B),
// so each tile is a single linear TMA bulk copy per K block.
//
// Roles (64 + 128*NS*NW threads): warp 0 = TMA producer, warp 1 = UMMA issuer (+ TMEM owner), then NS*NW epilogue
// warpgroups; S stage w in TMEM belongs to NW warpgroups (they split its 16-column chunks) and takes the column tiles n
// with n % NS == w (within a warpgroup each warp owns one TMEM lane quarter).  Narrow operands (<= 128 columns): NS =
// 3 or 4, NW = 1: 3-4 epilogue warps on every SM sub-partition (their MUFU / FMA dependency chains hide behind each
// other) and UMMA #1 runs NS-1 tiles ahead of the epilogue.  256-column operands: the O accumulator leaves room for
// NS = 2 stages only, so two warpgroups share each tile's epilogue (GramCfg).
//
// Tensor memory (512 columns): backward modes: O accumulator [0, min(KB,4)*64), S stages behind it at w*BN; forward
// modes: S stages at w*BN.  In the backward modes the epilogue overwrites the S stage it just consumed with W as packed
// bf16 (tcgen05.st) and UMMA #2 reads its A operand straight from TMEM -- W never touches shared memory.
//
// Work split: stream-K over (row block, column tile) pairs (PieceIter in plan.h), one accumulator copy per piece.
// Wide mode (KB = 8, 256 < D <= 512): A tile 128 KB resident, one B stage, O holds one 256-column half of dx per
// pass.  Row-sharded multi-GPU launches wait at their start for the peers' announcements (peer.cuh).
//
// Measured on B200 (tools/umma_rate.py): one thread can issue a tcgen05.mma about every 45 clk and an M=128, K=16
// instruction occupies the tensor pipe N/2 clk, so N >= 96 is needed to stay tensor-bound; the issue loop below is
// fully unrolled with precomputed descriptors for that reason.  Row tile = 128 rows (UMMA M), column tile = BN.
#pragma once
#include "peer.cuh"
#include "plan.h"
#include "ptx.cuh"
#include <type_traits>

namespace fb {

enum GramMode : int { NCE_FWD = 0, NCE_BWD = 1, TMP_FWD = 2, TMP_BWD = 3 };

// ---- tuning knobs (overridable with -D for tools/variant_bench.py experiments) ----
#ifndef FB_KB4_BN
#define FB_KB4_BN 96            // column tile for 256-wide operands (TMEM: 256 O columns + NS * BN <= 512)
#endif
#ifndef FB_KB4_NS
#define FB_KB4_NS 2
#endif
#ifndef FB_KB4_NB
#define FB_KB4_NB 3
#endif
#ifndef FB_KB4_NW
#define FB_KB4_NW 2             // epilogue warpgroups sharing one S stage (they split its 16-column chunks)
#endif
#ifndef FB_KB4_SHARE
#define FB_KB4_SHARE 1          // 256-column operands: 1 = ALL epilogue warpgroups work on every tile (see GramCfg::kShareAll)
#endif
#ifndef FB_KB4_SHARE_NW
#define FB_KB4_SHARE_NW 4       // epilogue warpgroups of the shared-stage configuration (4, or 6 = one 16-column chunk each at BN = 96)
#endif
#ifndef FB_KB2_NS_BWD
#define FB_KB2_NS_BWD 3         // narrow operands (<= 128 columns): S stages x warpgroups per stage of the backward passes
#endif
#ifndef FB_KB2_NW_BWD
#define FB_KB2_NW_BWD 1
#endif
#ifndef FB_KB2_NS_FWD
#define FB_KB2_NS_FWD 4         // ... and of the forward passes
#endif
#ifndef FB_KB2_NW_FWD
#define FB_KB2_NW_FWD 1
#endif
#ifndef FB_EPI_UNROLL
#define FB_EPI_UNROLL 1         // 2: two chunks of one thread's share of a tile are in flight together
#endif
#ifndef FB_EPI_HALF_CHUNKS
#define FB_EPI_HALF_CHUNKS 1    // see GramCfg::kHalfChunks
#endif
#ifndef FB_EPI_PIPE
#define FB_EPI_PIPE 0           // 1: one warpgroup per tile (narrow operands): software-pipelined tcgen05.ld, see FB_RUN_TILE.
                                // Measured (profiles/r2_variants.txt): row sums 88.5 -> 101.1 us, gradient pass 138.2 -> 142.3 us
                                // (the second register set spills at the 96 / 128 registers these kernels have): off
#endif
#ifndef FB_EPI_ROTATE
#define FB_EPI_ROTATE 1         // shared stages: rotate the chunk -> warpgroup assignment from tile to tile (see run_tile)
#endif
#ifndef FB_L2_HINTS
#define FB_L2_HINTS 0           // experiment: bit 0 = O-accumulator drain stores with L2 evict_last, bit 1 = feature loads /
#endif                          // gradient stores of the row kernels with evict_first (profiles/r2_l2_hints.txt)
#ifndef FB_POLY_PER8
#define FB_POLY_PER8 0          // columns (of every 8) whose exp2 runs as a polynomial on the FMA pipe instead of MUFU
#endif
#ifndef FB_POLY_FWD_PER8
#define FB_POLY_FWD_PER8 2      // same, for the row-sum pass only: that pass is MUFU-bound (67 % XU), and a degree-4
#endif                          // polynomial (rel. error 2.7e-6) for 2 of 8 columns took it from 93 to 84 us (0/1/2/3/4
                                // of 8: 93.3 / 89.1 / 84.0 / 84.8 / 87.1 us); the backward pass gains nothing from it
#ifndef FB_KB8_NS
#define FB_KB8_NS 2             // wide mode (8 K blocks): with one B stage the tiles are serial, so what counts is the
#endif                          // latency of one tile's epilogue: two warpgroups share it (NS x NW = 4 x 1 -> 2 x 2 took
#ifndef FB_KB8_NW               // the temporal launch of B = 4096, M = 4, D = 512 from 1118 to 899 us)
#define FB_KB8_NW 2
#endif
constexpr int kTmemCols = 512;

// Experiment builds only (tools/trace_gram.py, -DFB_TRACE=1): per-CTA clock64 stamps of the pipeline events of the first
// kTraceTiles column tiles: role 0 = TMA producer, 1 = UMMA issuer, 2 + g = epilogue warpgroup g.
#ifdef FB_TRACE
constexpr int kTraceTiles = 96, kTraceRoles = 6, kTraceTags = 4, kTraceBlocks = 160;
__device__ long long fb_trace_buf[kTraceBlocks * kTraceRoles * kTraceTiles * kTraceTags];
#define FB_TRACE_EV(role, n, tag)                                                                                     \
  do {                                                                                                                \
    if ((threadIdx.x & 31) == 0 && (n) < (uint32_t)kTraceTiles && blockIdx.x < kTraceBlocks)                          \
      fb_trace_buf[(((size_t)blockIdx.x * kTraceRoles + (role)) * kTraceTiles + (n)) * kTraceTags + (tag)] = clock64(); \
  } while (0)
#else
#define FB_TRACE_EV(role, n, tag) do {} while (0)
#endif

// Tile configuration as a function of the mode and the operand width in 64-element K blocks (see header comment).
// EL: tile precision.  0 = bf16 tiles (north_star's bf16 mode).  1 = split-bf16 tiles (the fp32 mode): every operand
// element x travels as hi = bf16(x) and lo = bf16(x - hi) -- 16 significant bits -- in two images (K blocks [0, KB/2) =
// hi, [KB/2, KB) = lo), and every product is three tensor-core passes hi*hi + hi*lo + lo*hi with fp32 accumulation.
// Why not kind::tf32: a 32-bit MN-major operand (UMMA #2's view of the B tile) is only accepted in the 32-byte-atom
// swizzle, in which a K-major operand (UMMA #1's view) faults, so no single shared-memory image can feed both GEMMs
// (tools/tf32_probe.py, profiles/r2_tf32_probe.txt), and two images of a D = 256 tile do not fit beside the 128 KB A tile.
// KB counts 128-byte K blocks, so the shared-memory / TMA side of a configuration depends on KB alone.
template <int MODE, int KB, int SEQ, int EL = 0>
struct GramCfg {
  static constexpr bool kBwd = (MODE == 1 || MODE == 3);
  static constexpr bool kTmp = (MODE >= 2);
  static constexpr int kEPB = EL ? 32 : 64;                                      // operand columns per K block (split: hi + lo)
  static_assert(EL == 0 || KB % 2 == 0, "split tiles: as many lo blocks as hi blocks");
  static constexpr bool kWide = (EL == 0 && KB > 4);                             // bf16, 256 < D <= 512: O in two passes
  static constexpr int BN = tile_bn(KB);                                         // column tile
  static constexpr int NS = KB <= 2 ? (kBwd ? FB_KB2_NS_BWD : FB_KB2_NS_FWD) : (KB == 4 ? FB_KB4_NS : (KB == 8 ? FB_KB8_NS : 4));  // S stages
  static constexpr int NB = KB <= 3 ? 5 : (KB == 4 ? FB_KB4_NB : 1);             // B-tile ring stages (smem budget)
  // kShareAll (256-column operands, two S stages): the pipeline trace (tools/trace_gram.py, profiles/r2_trace_temporal.txt)
  // showed the two stages' epilogues running one after the other, never together (each stage's chain UMMA #1 -> epilogue
  // -> UMMA #2 -> UMMA #1 is serial and the stages interleave), so warpgroups bound to a stage idle half of the time.
  // With kShareAll every warpgroup takes a share of EVERY tile: half the epilogue latency per tile, same issue work.
  static constexpr bool kShareAll = (KB == 4 && SEQ <= 16 && FB_KB4_SHARE != 0);
  static constexpr int NW = SEQ > 16 ? 1 : (KB == 4 ? (kShareAll ? FB_KB4_SHARE_NW : FB_KB4_NW) : (KB == 8 ? FB_KB8_NW : (KB <= 2 ? (kBwd ? FB_KB2_NW_BWD : FB_KB2_NW_FWD) : 1)));   // warpgroups per tile
  static constexpr int NG = kShareAll ? NW : NS * NW;                            // epilogue warpgroups in total
  static constexpr int CW = (NG >= 4 && (kTmp || NW > 1) && SEQ <= 16) ? 16 : 32;  // columns per tcgen05.ld (registers)
  // kHalfChunks (bf16 shared stages, 6 chunks of 16 columns over 4 warpgroups): every warpgroup takes one 16-column chunk
  // plus one 8-column half of chunks 4 / 5 -- 24 columns each -- instead of {2, 2, 1, 1} whole chunks: the epilogue latency
  // of a tile, which is on the critical path of each S stage's UMMA #1 -> epilogue -> UMMA #2 chain, drops from two chunk
  // times to one and a half.  The W of a split chunk k sits at packed columns [16k + 4, 16k + 12): both halves stay
  // inside the S columns their own warpgroup consumed, and the 8 columns of the K step remain contiguous.
  static constexpr bool kHalfChunks = kShareAll && EL == 0 && NW == 4 && CW == 16 && BN == 96 && SEQ <= 8 && FB_EPI_HALF_CHUNKS != 0;
  static constexpr int kThreads = 64 + 128 * NG;
  static_assert(NG <= 6, "partial-sum arrays are sized for <= 6 epilogue warpgroups");
  static constexpr int kOKB = kWide ? 4 : KB;        // K blocks of the O accumulator (wide mode: one half per pass)
  static constexpr int kON = kOKB * kEPB;            // columns of the O accumulator = UMMA #2 N
  static_assert(kON <= 256, "UMMA N");
  static_assert((kBwd ? kON : 0) + NS * BN <= kTmemCols, "TMEM budget");
};

template <int BN, int KB, int kNumBStages>
struct GramSmem {
  static constexpr uint32_t kABytes = KB * 128 * 128;           // [KB][128 rows][128 B]
  static constexpr uint32_t kBTile = KB * BN * 128;             // [KB][BN rows][128 B]
  static constexpr uint32_t kBStage = kBTile + 1024;            // tile + two per-column fp32 vectors
  static constexpr uint32_t kAOff = 0;
  static constexpr uint32_t kBOff = kAOff + kABytes;
  static constexpr uint32_t kBarOff = kBOff + kNumBStages * kBStage;
  static constexpr uint32_t kTotal = kBarOff + 8192;
  static constexpr uint32_t kDynamic = kTotal + 1024;           // slack for manual 1024-B alignment
  static_assert(2 * BN * 4 <= 1024, "column vectors fit the stage tail");
  static_assert((BN * 128) % 1024 == 0, "K blocks of a B tile stay 1024-B aligned");
};

struct GramBars {
  uint64_t a_full, a_empty;
  uint64_t b_full[8], b_empty[8];
  uint64_t s_full[4], s_empty[4];
  uint64_t w_full[4];
  uint64_t o_full, o_empty;
  uint32_t tmem_base;
  float red[4];
  float part_acc[5][128];     // per-row partials of epilogue warpgroups 1.., folded into warpgroup 0 at the end of an item
  float part_hinge[5][128];
  int32_t part_cnt[5][128];
};
static_assert(sizeof(GramBars) <= 8192, "barrier block");

// ---------------------------------------------------------------------------------------------------------
// work-item decoding (everything a role needs about one 128-row block; no arrays, lives in registers)
// ---------------------------------------------------------------------------------------------------------
struct Item {
  const uint8_t* a_src;        // first K block of the A tile
  const uint8_t* b_src0;       // column side 0 / 1 operand base (row 0, K block 0)
  const uint8_t* b_src1;
  uint64_t kstride;            // bytes between K blocks of the operand arrays
  const float* cv0_0;          // per-column vector #0 of side 0 / 1 (NCE: 1/rowsum, TMP: squared norms)
  const float* cv0_1;
  const float* cv1;            // per-column vector #1 (TMP: m_JJ); same for both sides
  int ct_begin, ct_end;        // global column tile range
  int ntc;                     // column tiles per side
  int row0;                    // first row (within side / tensor) of the row tile
  int side;                    // NCE: which half of z the rows come from
  int ncol_valid;              // valid columns per side (b for NCE, B for TMP)
  int row_lo, row_hi;          // owned rows [lo, hi) within the side
  int q, s, c;                 // problem / position / (split or call)
};

template <int MODE>
__device__ __forceinline__ int gram_num_items(const Plan& p, const ProbSel& sel) {
  if (MODE == NCE_FWD || MODE == NCE_BWD) {
    const int t0 = p.seq0 / kTileM, t1 = (p.seq1 + kTileM - 1) / kTileM;
    return sel.n * p.S * 2 * (t1 - t0);
  } else {
    const int t0 = (p.seq0 * p.Sp) / kTileM, t1 = (p.seq1 * p.Sp + kTileM - 1) / kTileM;
    return p.nT * (t1 - t0) * ((MODE == TMP_BWD && p.wide) ? 2 : 1);           // wide mode: one item per output half
  }
}

template <int MODE, int BN>
__device__ __forceinline__ void gram_decode(const Plan& p, const ProbSel& sel, const uint8_t* ws, int it, Item& x) {
  if (MODE == NCE_FWD || MODE == NCE_BWD) {
    const int t0 = p.seq0 / kTileM, t1 = (p.seq1 + kTileM - 1) / kTileM, nrt = t1 - t0;
    int r = it;
    const int rt = t0 + r % nrt; r /= nrt;
    const int side = r % 2; r /= 2;
    const int s = r % p.S; r /= p.S;
    const int q = sel.idx[r];
    const ProbDesc& pr = p.probs[q];
    x.kstride = (uint64_t)p.S * p.bpad * 128;
    x.b_src0 = ws + p.ops[pr.opA].off + (uint64_t)s * p.bpad * 128;
    x.b_src1 = ws + p.ops[pr.opB].off + (uint64_t)s * p.bpad * 128;
    x.a_src = (side ? x.b_src1 : x.b_src0) + (uint64_t)rt * kTileM * 128;
    const float* rinv = reinterpret_cast<const float*>(ws + p.rinv_off) + ((uint64_t)(q * p.S + s) * 2) * p.bpad;
    x.cv0_0 = rinv; x.cv0_1 = rinv + p.bpad; x.cv1 = rinv;
    x.ntc = (p.b + BN - 1) / BN;
    x.ct_begin = 0;
    x.ct_end = 2 * x.ntc;
    x.row0 = rt * kTileM; x.side = side; x.ncol_valid = p.b;
    x.row_lo = p.seq0; x.row_hi = p.seq1;
    x.q = q; x.s = s; x.c = 0;
  } else {
    const int t0 = (p.seq0 * p.Sp) / kTileM, t1 = (p.seq1 * p.Sp + kTileM - 1) / kTileM, nrt = t1 - t0;
    int r = it, half = 0;
    if (MODE == TMP_BWD && p.wide) { half = r & 1; r >>= 1; }
    const int rt = t0 + r % nrt;
    const int c = r / nrt;
    x.kstride = (uint64_t)p.Bpad * 128;
    x.b_src0 = x.b_src1 = ws + p.xt_off + (uint64_t)c * p.kbFull * p.Bpad * 128;
    x.a_src = x.b_src0 + (uint64_t)rt * kTileM * 128;
    x.cv0_0 = x.cv0_1 = reinterpret_cast<const float*>(ws + p.sq_off) + (uint64_t)c * p.Bpad;
    x.cv1 = reinterpret_cast<const float*>(ws + p.mintra_off) + (uint64_t)c * p.Bpad;
    x.ntc = (p.Bt + BN - 1) / BN;                    // rows and columns live in the temporal row space (plan.h: Sp, Bt)
    x.ct_begin = 0; x.ct_end = x.ntc;
    x.row0 = rt * kTileM; x.side = 0; x.ncol_valid = p.Bt;
    x.row_lo = p.seq0 * p.Sp; x.row_hi = p.seq1 * p.Sp;
    x.q = half; x.s = 0; x.c = c;                   // q: which 256-column half of dx this item accumulates (wide mode)
  }
}

// Stream-K work split: the (row block, column tile) pairs of a launch are numbered consecutively and every CTA takes an
// equal contiguous share, so a row block may be cut into a primary piece (starts at tile 0) and secondary pieces
// (handled by the following CTAs): one when a share is at least a row block long, up to kMaxPieces - 1 when a row shard
// leaves fewer row blocks than SMs.  All SMs stay busy until the end of the launch (with whole row blocks 256 equal
// items on 148 SMs ran 2 rounds for 1.73 rounds of work).  Piece k writes to its own copy of the accumulators
// (dz / dx / rho / cnt / row sums + k * delta) and finalize adds them in piece order, so the result is deterministic.
template <int MODE, int BN>
__device__ __forceinline__ int gram_tiles_per_item(const Plan& p) {
  return (MODE == NCE_FWD || MODE == NCE_BWD) ? 2 * ((p.b + BN - 1) / BN) : (p.Bt + BN - 1) / BN;
}

__device__ __forceinline__ float4 lds128(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ float lds32(uint32_t addr) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
  return v;
}

// ---------------------------------------------------------------------------------------------------------
// the kernel
// ---------------------------------------------------------------------------------------------------------
// Exchange lane bit LB with register-index bit RB of a 32 x 32 block held as v[column] in lane = row: afterwards lane L,
// register c holds the element whose (row, column) are (L, c) with those two bits swapped.  16 shuffles.
template <int LB, int RB>
__device__ __forceinline__ void xchg_lane_reg_bit(float (&v)[32], int lane) {
  const bool hi = ((lane >> LB) & 1) != 0;
#pragma unroll
  for (int a = 0; a < 32; ++a) {
    if (a & (1 << RB)) continue;
    constexpr int kB = 1 << RB;
    const float send = hi ? v[a] : v[a | kB];
    const float recv = __shfl_xor_sync(0xffffffffu, send, 1 << LB);
    if (hi) v[a] = recv; else v[a | kB] = recv;
  }
}

template <int MODE, int KB, int SEQ, int EL>
__global__ void __launch_bounds__((GramCfg<MODE, KB, SEQ, EL>::kThreads), 1)
gram_kernel(const __grid_constant__ Plan p, const __grid_constant__ ProbSel sel, uint8_t* __restrict__ ws) {
  using G = GramCfg<MODE, KB, SEQ, EL>;
  constexpr int BN = G::BN, NS = G::NS, NB = G::NB, CW = G::CW, NW = G::NW, NG = G::NG;
  using L = GramSmem<BN, KB, NB>;
  constexpr bool kIsNce = (MODE == NCE_FWD || MODE == NCE_BWD);
  constexpr bool kBwd = (MODE == NCE_BWD || MODE == TMP_BWD);
  constexpr bool kColVec = (MODE != NCE_FWD);
  constexpr int kEpiThreads = 128 * NG;
  // TMEM columns between the W of consecutive UMMA #2 K steps (16 bf16 = 8 packed columns).  When two warpgroups share a
  // stage each W chunk stays inside the 16 S columns its own warpgroup consumed, so nobody overwrites columns the
  // other warpgroup may not have read yet.  Split tiles: chunk ch keeps its own CW columns: [hi pairs | lo pairs].
  constexpr bool kSplit = (EL == 1);
  constexpr bool kWide = G::kWide;
  constexpr int kKH = kSplit ? KB / 2 : KB;          // K blocks of one image (split: hi and lo images of kKH blocks each)
  constexpr int kWStep = (NW > 1) ? 16 : 8;
  static_assert(NW == 1 || CW == 16, "shared stages use 16-column chunks");
  constexpr int kON = G::kON;                        // UMMA #2 N = padded operand width (wide mode: one half of it)
  constexpr uint32_t kOCol = 0;                      // TMEM: O accumulator at [0, kON) (backward modes only)
  constexpr uint32_t kSCol = kBwd ? kON : 0;         // TMEM: S stage w at kSCol + w * BN
  static_assert(BN % CW == 0 && (SEQ == 0 || CW % (SEQ > 0 ? SEQ : 1) == 0), "column tile / chunk / sequence");
  static_assert(L::kDynamic <= 232448, "shared memory budget");

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  GramBars* bars = reinterpret_cast<GramBars*>(smem + L::kBarOff);
  // warp-uniform role index (the compiler must be able to prove uniformity, see the issue-loop note below)
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;

  pdl_launch_dependents();      // the next launch may be scheduled; it waits for this grid's completion itself
  if (threadIdx.x == 0) FB_TRACE_EV(0, 0u, 2);       // kernel entry
  if (threadIdx.x == 0) {
    mbar_init(&bars->a_full, 1);
    mbar_init(&bars->a_empty, 1);
    for (int i = 0; i < NB; ++i) {
      mbar_init(&bars->b_full[i], 1);
      mbar_init(&bars->b_empty[i], kColVec ? 1 + 4 * NW : 1);   // UMMA commit (+ one elected lane per epilogue warp)
    }
    for (int i = 0; i < NS; ++i) {
      mbar_init(&bars->s_full[i], 1);
      mbar_init(&bars->s_empty[i], 128 * NW);
      mbar_init(&bars->w_full[i], 128 * NW);
    }
    mbar_init(&bars->o_full, 1);
    mbar_init(&bars->o_empty, kEpiThreads);
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(&bars->tmem_base, kTmemCols);
    tmem_relinquish();
  }
  // everything above overlapped the tail of the previous launch; from here on its results are needed
  pdl_wait();
  // row-sharded path: the launch before this one stored operands / row sums into the peers' workspaces -- tell them
  if (sel.ann_world > 0 && blockIdx.x == 0) peer_announce_epoch(p, sel.peer_ws, sel.ann_world, sel.ann_rank);
  if (sel.peer_wait > 0) {
    // row-sharded path: the operands / row sums this launch reads were stored by the peers; wait for their
    // announcements (block 0 of the launch after their producing launch sent them), then order the TMA reads behind the wait
    peer_wait(ws, p, sel.peer_wait, -1);
    asm volatile("fence.proxy.async;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = __shfl_sync(0xffffffffu, bars->tmem_base, 0);
  const int n_items = gram_num_items<MODE>(p, sel);
  if (threadIdx.x == 0) FB_TRACE_EV(0, 0u, 3);       // set-up done (barriers, TMEM, peer wait)

  // Issue-loop note (measured, tools/umma_rate.py / tools/tma_rate.py): tcgen05.mma, tcgen05.commit and cp.async.bulk
  // take their operands from uniform registers.  If ptxas cannot prove an operand warp-uniform it wraps EVERY such
  // instruction in an ELECT / R2UR.BROADCAST / BRA.U.ANY waterfall (~80-270 clk per instruction instead of ~10), which
  // made the single issuing lane -- not the tensor pipe -- the bound of earlier versions.  Therefore the producer
  // and issuer warps run their loops with all 32 lanes on warp-uniform values and only the asm itself sits under
  // elect_one().

  if (warp == 0) {
    // =============================== TMA producer ===============================
    {
      uint32_t nb = 0, ni = 0;
      PieceIter pieces((int)blockIdx.x, (int)gridDim.x, n_items, gram_tiles_per_item<MODE, BN>(p),
                       (kIsNce ? p.sk_nce : p.sk_tmp) != 0);
      for (int it, pt0, pt1, pk, npi; pieces.next(it, pt0, pt1, pk, npi); ++ni) {
        Item x;
        gram_decode<MODE, BN>(p, sel, ws, it, x);
        x.ct_begin = pt0; x.ct_end = pt1;
        mbar_wait_warp(&bars->a_empty, (ni & 1) ^ 1);
        FB_TRACE_EV(0, nb, 1);                        // A tile of the piece that starts with column tile nb requested
        if (elect_one()) {
          mbar_arrive_expect_tx(&bars->a_full, L::kABytes);
#pragma unroll
          for (int kb = 0; kb < KB; ++kb)
            tma_load_1d(smem + L::kAOff + kb * 16384, x.a_src + kb * x.kstride, 16384, &bars->a_full);
        }
        __syncwarp();
        for (int ct = x.ct_begin; ct < x.ct_end; ++ct, ++nb) {
          const uint32_t st = nb % NB;
          const int cs = ct >= x.ntc ? 1 : 0, tc = ct - cs * x.ntc;
          const uint8_t* src = (cs ? x.b_src1 : x.b_src0) + (uint64_t)tc * BN * 128;
          mbar_wait_warp(&bars->b_empty[st], ((nb / NB) & 1) ^ 1);
          FB_TRACE_EV(0, nb, 0);
          if (elect_one()) {
            uint8_t* dst = smem + L::kBOff + st * L::kBStage;
            constexpr uint32_t bytes = L::kBTile + (kColVec ? (kIsNce ? 1 : 2) * BN * 4 : 0);
            mbar_arrive_expect_tx(&bars->b_full[st], bytes);
#pragma unroll
            for (int kb = 0; kb < KB; ++kb)
              tma_load_1d(dst + kb * (BN * 128), src + kb * x.kstride, BN * 128, &bars->b_full[st]);
            if (kColVec) {
              tma_load_1d(dst + L::kBTile, (cs ? x.cv0_1 : x.cv0_0) + tc * BN, BN * 4, &bars->b_full[st]);
              if (!kIsNce) tma_load_1d(dst + L::kBTile + BN * 4, x.cv1 + tc * BN, BN * 4, &bars->b_full[st]);
            }
          }
          __syncwarp();
        }
      }
    }
  } else if (warp == 1) {
    // =============================== UMMA issuer ===============================
    {
      constexpr uint32_t idesc1 = umma_idesc(UMMA_BF16, 128, BN, 0, 0);     // S = A(K-major) * B(K-major)^T
      constexpr uint32_t idesc2 = umma_idesc(UMMA_BF16, 128, kON, 0, 1);    // O += W(TMEM) * B(MN-major)
      const uint64_t da0 = umma_smem_desc(smem_u32(smem + L::kAOff), 16, 1024);
      const uint64_t db0 = umma_smem_desc(smem_u32(smem + L::kBOff), 16, 1024);            // K-major view
      const uint64_t dm0 = umma_smem_desc(smem_u32(smem + L::kBOff), BN * 128, 1024);      // MN-major view
      // UMMA #1 of one column tile: K steps of 16 over the kKH blocks of an image; split tiles: three passes
      // hi*hi + hi*lo + lo*hi (descriptor offsets of the lo images: kKH blocks further)
      auto issue_gram = [&](uint32_t d, uint64_t db) {
        constexpr int kPasses = kSplit ? 3 : 1;
#pragma unroll
        for (int ps = 0; ps < kPasses; ++ps) {
          const uint64_t ao = (uint64_t)(((ps == 2) ? kKH * 16384 : 0) >> 4);
          const uint64_t bo = (uint64_t)(((ps == 1) ? kKH * (BN * 128) : 0) >> 4);
#pragma unroll
          for (int k = 0; k < kKH * 4; ++k)
            umma_bf16(d, da0 + ao + (uint64_t)(((k >> 2) * 16384 + (k & 3) * 32) >> 4),
                      db + bo + (uint64_t)(((k >> 2) * (BN * 128) + (k & 3) * 32) >> 4), idesc1, (ps | k) > 0);
        }
      };
      uint32_t nb = 0, ni = 0;
      PieceIter pieces((int)blockIdx.x, (int)gridDim.x, n_items, gram_tiles_per_item<MODE, BN>(p),
                       (kIsNce ? p.sk_nce : p.sk_tmp) != 0);
      for (int it, pt0, pt1, pk, npi; pieces.next(it, pt0, pt1, pk, npi); ++ni) {
        Item x;
        gram_decode<MODE, BN>(p, sel, ws, it, x);
        x.ct_begin = pt0; x.ct_end = pt1;
        mbar_wait_warp(&bars->a_full, ni & 1);
        const int ntiles = x.ct_end - x.ct_begin;
        if (!kBwd) {
          for (int t = 0; t < ntiles; ++t) {
            // ---- UMMA #1 of tile t into S stage n % NS (free once the epilogue has drained it)
            const uint32_t n = nb + t, st = n % NB, ss = n % NS;
            mbar_wait_warp(&bars->b_full[st], (n / NB) & 1);
            mbar_wait_warp(&bars->s_empty[ss], ((n / NS) & 1) ^ 1);
            tc_fence_after();
            FB_TRACE_EV(1, n, 0);
            if (elect_one()) {
              issue_gram(tmem + kSCol + ss * BN, db0 + (uint64_t)(st * (L::kBStage >> 4)));
              umma_commit(&bars->s_full[ss]);
              umma_commit(&bars->b_empty[st]);
            }
            __syncwarp();
            FB_TRACE_EV(1, n, 1);
          }
        } else {
          // Out-of-order issue: UMMA #2 of the oldest tile whose W is ready, else UMMA #1 of the next tile whose B
          // tile has landed and whose S stage is free (UMMA #2 of tile t1 - NS already issued).  Neither wait may
          // block the other: the B ring is too short (shared memory) to hide a blocked issuer behind prefetch.
          int t1 = 0, t2 = 0;
          while (t2 < ntiles) {
            bool did = false;
            if (t2 < t1) {
              const uint32_t n = nb + t2, st = n % NB, ss = n % NS;
              if (t2 == 0) mbar_wait_warp(&bars->o_empty, (ni & 1) ^ 1);
              const bool ready = mbar_try_wait_warp(&bars->w_full[ss], (n / NS) & 1);
              if (ready) {
                tc_fence_after();
                FB_TRACE_EV(1, n, 2);
                if (elect_one()) {
                  // MN-major view of the B tile; wide mode: the K blocks of this item's output half
                  const uint64_t dm = dm0 + (uint64_t)(st * (L::kBStage >> 4)) +
                                      (uint64_t)((kWide ? x.q * 4 * (BN * 128) : 0) >> 4);
                  const uint32_t a = tmem + kSCol + ss * BN;    // W: packed bf16 over the consumed S stage
                  const uint32_t acc = (t2 > 0) ? 1u : 0u;
                  if constexpr (!kSplit) {
#pragma unroll
                    for (int k = 0; k < BN / 16; ++k)
                      umma_bf16_ts(tmem + kOCol, a + k * kWStep + ((G::kHalfChunks && k >= NW) ? 4 : 0),
                                   dm + (uint64_t)((k * 2048) >> 4), idesc2, k > 0 ? 1u : acc);
                  } else {
                    // split tiles: chunk c of CW columns holds [W_hi pairs | W_lo pairs]; O += Wh Xh + Wh Xl + Wl Xh
                    constexpr uint64_t lo_img = (uint64_t)((kKH * (BN * 128)) >> 4);
#pragma unroll
                    for (int k = 0; k < BN / 16; ++k) {
                      const uint32_t ah = a + (k * 16 / CW) * CW + (k % (CW / 16)) * 8, al = ah + CW / 2;
                      const uint64_t dk = dm + (uint64_t)((k * 2048) >> 4);
                      umma_bf16_ts(tmem + kOCol, ah, dk, idesc2, k > 0 ? 1u : acc);
                      umma_bf16_ts(tmem + kOCol, ah, dk + lo_img, idesc2, 1u);
                      umma_bf16_ts(tmem + kOCol, al, dk, idesc2, 1u);
                    }
                  }
                  umma_commit(&bars->b_empty[st]);
                }
                __syncwarp();
                FB_TRACE_EV(1, n, 3);
                ++t2;
                did = true;
              }
            }
            if (t1 < ntiles && t1 - t2 < NS) {
              const uint32_t n = nb + t1, st = n % NB, ss = n % NS;
              const bool ready = mbar_try_wait_warp(&bars->b_full[st], (n / NB) & 1);
              if (ready) {
                tc_fence_after();
                FB_TRACE_EV(1, n, 0);
                if (elect_one()) {
                  issue_gram(tmem + kSCol + ss * BN, db0 + (uint64_t)(st * (L::kBStage >> 4)));
                  umma_commit(&bars->s_full[ss]);
                  // The A tile is only read by UMMA #1: hand it back as soon as the LAST one of the piece has executed,
                  // so the next piece's A load overlaps this piece's last epilogue + UMMA #2 (it used to start after
                  // them: ~7 700 clk lost per piece boundary of the temporal launch, profiles/r2_timeline_r1.txt).
                  if (t1 + 1 == ntiles) umma_commit(&bars->a_empty);
                }
                __syncwarp();
                FB_TRACE_EV(1, n, 1);
                ++t1;
                did = true;
              }
            }
            if (!did) __nanosleep(20);
          }
        }
        nb += ntiles;
        if (elect_one()) {
          if (!kBwd || ntiles == 0) umma_commit(&bars->a_empty);   // backward: went out behind the last UMMA #1 above
          if (kBwd) umma_commit(&bars->o_full);
        }
        __syncwarp();
      }
      // Commits complete in issue order.  Forward passes: once the last a_empty has landed no arrival is still in flight.
      // Backward passes: the last commit is o_full, which the epilogue warps wait for before the final __syncthreads.
      if (ni > 0) mbar_wait_warp(&bars->a_empty, (ni - 1) & 1);
    }
  } else {
    // =============================== epilogue warps ===============================
    constexpr bool kShareAll = G::kShareAll;          // every warpgroup works on every tile (stage = tile index % NS)
    const int wgi = (warp - 2) >> 2;                  // epilogue warpgroup index
    const int wg0 = kShareAll ? 0 : wgi / NW;         // S stage this warpgroup is bound to (unless kShareAll)
    const int sub = kShareAll ? wgi : wgi % NW;       // which of a tile's chunks it takes (chunk % NW == sub)
    const int quarter = warp & 3;                     // TMEM lane quarter this warp may access
    const int trow = quarter * 32 + lane;             // row within the tile == TMEM lane
    const uint32_t tlane = (uint32_t)(quarter * 32) << 16;
    constexpr int SQ = SEQ > 0 ? SEQ : 1;             // lanes / columns per (padded) sequence: plan.h Sp
    const int Sr = kIsNce ? 1 : p.S;                  // real sequence length; positions >= Sr of a sequence are phantoms
    const bool padseq = !kIsNce && Sr != SQ;
    const float inv_cnt = 1.f / (float)(Sr * Sr);
    const float coef1 = -1.f / ((float)p.b * (float)(p.b - 1) * (float)(Sr * Sr));
    const float coef2 = 2.f * coef1;
    const float margin = p.margin;
    const uint32_t cv_base = smem_u32(smem + L::kBOff + L::kBTile);
    const uint32_t s_addr0 = tmem + tlane + kSCol;
    uint32_t nb = 0, ni = 0;
    PieceIter pieces((int)blockIdx.x, (int)gridDim.x, n_items, gram_tiles_per_item<MODE, BN>(p),
                       (kIsNce ? p.sk_nce : p.sk_tmp) != 0);
    for (int it, pt0, pt1, pk, npi; pieces.next(it, pt0, pt1, pk, npi); ++ni) {
      Item x;
      gram_decode<MODE, BN>(p, sel, ws, it, x);
      x.ct_begin = pt0; x.ct_end = pt1;
      const bool second = pk > 0;                     // secondary piece of a split row block
      const int row0 = x.row0, ncol_valid = x.ncol_valid, side = x.side, ntc = x.ntc, ct_begin = x.ct_begin;
      const int row = row0 + trow;                    // row within side (NCE: sequence index k) / tensor (TMP: i)
      const bool rowpad = padseq && (row & (SQ - 1)) >= Sr;            // phantom row of a padded sequence
      const bool row_ok = row < ncol_valid && row >= x.row_lo && row < x.row_hi && !rowpad;
      float rowacc = 0.f;                             // NCE_FWD: row sum; TMP: rho_i (both accumulate in racc, packed pairs)
      float racc[4] = {0.f, 0.f, 0.f, 0.f};
      float posg = 0.f;                               // NCE_FWD: G_{k,p(k)} as this row's tile computed it
      bool pos_seen = false;
      float ck = 0.f, n_i = 0.f, mim = -1e30f, hinge_acc = 0.f;
      int cnt_i = 0;
      if (MODE == NCE_BWD && row_ok) ck = (side ? x.cv0_1 : x.cv0_0)[row];
      if (!kIsNce && row_ok) { n_i = x.cv0_0[row]; mim = x.cv1[row]; }            // m_II
      const int seq_i = row / SQ;
      const int ntiles = x.ct_end - ct_begin;
      // first tile of this item that belongs to this warpgroup: (nb + t) % NS == wg0; kShareAll: every tile
      for (int t = kShareAll ? 0 : (int)((wg0 + NS - nb % NS) % NS); t < ntiles; t += (kShareAll ? 1 : NS)) {
        const uint32_t n = nb + t, st = n % NB;
        const uint32_t sphase = (n / NS) & 1;
        const int wg = kShareAll ? (int)(n % NS) : wg0;     // S stage of this tile
        const uint32_t s_addr = s_addr0 + wg * BN;
        const int ct = ct_begin + t;
        const int cs = ct >= ntc ? 1 : 0, tc = ct - cs * ntc;
        const int col0 = tc * BN;                     // first column (within side) of this tile
        const uint32_t cv = cv_base + st * L::kBStage;
        FB_TRACE_EV(2 + wgi, n, 0);
        if (kColVec) mbar_wait(&bars->b_full[st], (n / NB) & 1);
        mbar_wait(&bars->s_full[wg], sphase);
        tc_fence_after();
        FB_TRACE_EV(2 + wgi, n, 1);
        const bool tail = col0 + BN > ncol_valid;
        // columns to drop: j == k (same side) always; in the backward pass also the positive p(k) (other side,
        // same sequence index), whose contribution the finalize kernel adds in fp32
        const bool overlap = col0 < row0 + kTileM && col0 + BN > row0;
        const bool diag = kIsNce ? (overlap && (MODE == NCE_BWD || cs == side)) : overlap;
        // the positive p(k): other side, same sequence index.  The row-sum pass hands the logit it saw to nce_lse /
        // finalize, so that ln(sum_j e^{s_kj}) - s_{k,p(k)} cancels to the last bit where the positive dominates the row
        // (tensor-core accumulation is not round-to-nearest: a separately computed fp32 dot product differs by ~1e-5)
        const bool ptile = (MODE == NCE_FWD) && overlap && cs != side;
        // One chunk of CW columns.  EDGE = the tile touches the (block) diagonal, the padded tail of the columns, or the
        // sequences are padded (seq_len not a power of two): only those tiles pay for the per-pair validity logic;
        // interior tiles -- nearly all of them -- run the short path.
        // coff: first column of the chunk within the tile; its width W comes with the register array; wcol: TMEM column
        // (relative to the S stage) where the chunk's W goes
        auto chunk_body = [&](auto edge_c, int coff, auto& v, int wcol) {
          constexpr bool EDGE = decltype(edge_c)::value;
          constexpr int W = (int)(sizeof(v) / sizeof(float));
          const int cbase = col0 + coff;              // column (within side) of v[0]
          if constexpr (kIsNce) {
            if (EDGE && ptile) {
#pragma unroll
              for (int j = 0; j < W; ++j)
                if (cbase + j == row) { posg = v[j]; pos_seen = true; }
            }
            // ---------------- InfoNCE: E = 2^G (logits arrive pre-scaled to the log2 domain)
#pragma unroll
            for (int j = 0; j < W; ++j)
              v[j] = ((j & 7) >= 8 - (MODE == NCE_FWD ? FB_POLY_FWD_PER8 : FB_POLY_PER8)) ? ex2_poly(v[j]) : ex2_approx(v[j]);
            if (MODE == NCE_BWD) {
              // W_kj = E_kj (1/r_k + 1/r_j)  ==  P_kj + P_jk  
#pragma unroll
              for (int j = 0; j < W; j += 4) {
                const float4 cj = lds128(cv + (coff + j) * 4);
                float c0, c1, c2, c3;
                add2(c0, c1, cj.x, cj.y, ck, ck);
                add2(c2, c3, cj.z, cj.w, ck, ck);
                mul2(v[j], v[j + 1], v[j], v[j + 1], c0, c1);
                mul2(v[j + 2], v[j + 3], v[j + 2], v[j + 3], c2, c3);
              }
            }
            if (EDGE) {
#pragma unroll
              for (int j = 0; j < W; ++j) {
                const int col = cbase + j;
                if ((diag && col == row) || col >= ncol_valid) v[j] = 0.f;
              }
            }
            if (MODE == NCE_FWD) {
#pragma unroll
              for (int j = 0; j < W; j += 4) {
                add2(racc[0], racc[1], racc[0], racc[1], v[j], v[j + 1]);
                add2(racc[2], racc[3], racc[2], racc[3], v[j + 2], v[j + 3]);
              }
            }
          } else {
            // ---------------- temporal: delta_ij, S x S block means, hinge, r_ij
            float nj[W];
#pragma unroll
            for (int j = 0; j < W; j += 4) {
              const float4 t4 = lds128(cv + (coff + j) * 4);
              add2(nj[j], nj[j + 1], t4.x, t4.y, n_i, n_i);
              add2(nj[j + 2], nj[j + 3], t4.z, t4.w, n_i, n_i);
            }
#pragma unroll
            for (int g0 = 0; g0 < W; g0 += SQ) {
              float gsum = 0.f;
#pragma unroll
              for (int j = 0; j < SQ; j += 2) {
                // cdist mm form; the floor keeps 1/delta finite for coincident rows (their r_ij (x_i - x_j) is 0)
                float d0, d1;
                fma2(d0, d1, v[g0 + j], v[g0 + j + 1], -2.f, -2.f, nj[g0 + j], nj[g0 + j + 1]);
                d0 = fmaxf(d0, 1e-12f); d1 = fmaxf(d1, 1e-12f);
                float r0 = rsqrt_approx(d0), r1 = rsqrt_approx(d1);   // 1 / delta
                if (EDGE && padseq) {
                  // sequence length not a power of two: columns at positions >= Sr are phantoms (no distance, no weight)
                  if (j >= Sr) { r0 = 0.f; d0 = 0.f; }
                  if (j + 1 >= Sr) { r1 = 0.f; d1 = 0.f; }
                }
                v[g0 + j] = r0; v[g0 + j + 1] = r1;
                gsum = fmaf(d0, r0, gsum);                              // delta = d2 / delta
                gsum = fmaf(d1, r1, gsum);
              }
              if (EDGE && rowpad) gsum = 0.f;                           // phantom rows add nothing to the block sums
#pragma unroll
              for (int o = 1; o < SQ; o <<= 1) gsum += __shfl_xor_sync(0xffffffffu, gsum, o);
              const float mm = fmaf(gsum, inv_cnt, -margin);            // m_IJ - margin
              const float h = mim - mm;                                 // hinge argument (mim = m_II; -1e30 on rows that are not ok)
              const float hj = lds32(cv + (BN + coff + g0) * 4) - mm;   // ... of the transposed pair: m_JJ + margin - m_IJ
              bool a_ij = h >= 0.f, a_ji = hj >= 0.f;                   // active at equality
              if (EDGE) {
                const int colg = cbase + g0;
                bool pair_ok = row_ok;
                if (tail) pair_ok = pair_ok && colg < ncol_valid;
                if (diag) pair_ok = pair_ok && (colg / SQ) != seq_i;    // the block diagonal is done exactly elsewhere
                a_ij = a_ij && pair_ok;
                a_ji = a_ji && pair_ok;
              }
              // branch-free on purpose: a branch per group ends the basic block, and the groups of a chunk then run one
              // after the other instead of interleaved (coef2 == 2 coef1 exactly, so the sum of the two selects is exact)
              hinge_acc += a_ij ? h : 0.f;
              cnt_i += a_ij ? 1 : 0;
              const float coef = (a_ij ? coef1 : 0.f) + (a_ji ? coef1 : 0.f);
#pragma unroll
              for (int j = 0; j < SQ; j += 2) {
                mul2(v[g0 + j], v[g0 + j + 1], v[g0 + j], v[g0 + j + 1], coef, coef);
                add2(racc[j & 2], racc[(j & 2) + 1], racc[j & 2], racc[(j & 2) + 1], v[g0 + j], v[g0 + j + 1]);
              }
            }
          }
          if (kBwd) {
            if (kSplit) {
              // ---------------- W chunk, split: [hi pairs | lo pairs] over the CW columns this chunk came from
              uint32_t wv[W];
#pragma unroll
              for (int j = 0; j < W / 2; ++j) {
                const uint32_t hi = pack_bf16x2(v[2 * j], v[2 * j + 1]);
                wv[j] = hi;
                wv[W / 2 + j] = pack_bf16x2(v[2 * j] - __uint_as_float(hi << 16), v[2 * j + 1] - __uint_as_float(hi & 0xffff0000u));
              }
              if constexpr (W >= 16) tmem_st_full<W>(s_addr + wcol, wv);
            } else {
              // ---------------- W chunk: packed bf16 over the S columns this thread has already consumed
              uint32_t pk[W / 2];
#pragma unroll
              for (int j = 0; j < W / 2; ++j) pk[j] = pack_bf16x2(v[2 * j], v[2 * j + 1]);
              tmem_st_packed<W>(s_addr + wcol, pk);
            }
          }
        };
        // Shared stages whose chunk count is not a multiple of the warpgroup count (96 columns = 6 chunks over 4
        // warpgroups): rotate the assignment from tile to tile so that every warpgroup does 3 chunks per two tiles
        // instead of {2, 2, 1, 1} on every tile (the warpgroups move on to the next tile independently).
        const int sub_t = (kShareAll && !G::kHalfChunks && FB_EPI_ROTATE && (BN / CW) % NW != 0) ? ((sub + (int)(n & 1) * (NW / 2)) & (NW - 1)) : sub;
        // FB_EPI_UNROLL == 2: two chunks of this thread in flight together: twice the independent work behind each
        // tcgen05.ld / MUFU / shuffle latency
#define FB_WCOL(ch) ((ch) * (kSplit ? CW : (NW > 1 ? CW : CW / 2)))
#define FB_RUN_TILE(EDGE_TAG)                                                        \
  {                                                                                  \
    if constexpr (G::kHalfChunks) {                                                  \
      /* chunk `sub` whole + one half of chunk 4 / 5 (24 columns per warpgroup) */   \
      const int hoff = (NW + (sub >> 1)) * CW + (sub & 1) * (CW / 2);                \
      float v0[CW], v1[CW / 2];                                                      \
      tmem_ld_chunk<CW>(s_addr + sub * CW, v0);                                      \
      tmem_ld_chunk<CW / 2>(s_addr + hoff, v1);                                      \
      tmem_ld_wait();                                                                \
      chunk_body(EDGE_TAG, sub * CW, v0, sub * CW);                                  \
      chunk_body(EDGE_TAG, hoff, v1, (NW + (sub >> 1)) * CW + 4 + (sub & 1) * 4);    \
    } else if constexpr (NW == 1 && FB_EPI_PIPE != 0) {                              \
      /* one warpgroup per tile: the tcgen05.ld of chunk c + 1 is in flight while chunk c is computed */ \
      constexpr int kCh = BN / CW;                                                   \
      float va[CW], vb[CW];                                                          \
      tmem_ld_chunk<CW>(s_addr, va);                                                 \
      _Pragma("unroll") for (int ch = 0; ch < kCh; ch += 2) {                        \
        tmem_ld_wait();                                                              \
        if (ch + 1 < kCh) tmem_ld_chunk<CW>(s_addr + (ch + 1) * CW, vb);             \
        chunk_body(EDGE_TAG, ch * CW, va, FB_WCOL(ch));                              \
        if (ch + 1 < kCh) {                                                          \
          tmem_ld_wait();                                                            \
          if (ch + 2 < kCh) tmem_ld_chunk<CW>(s_addr + (ch + 2) * CW, va);           \
          chunk_body(EDGE_TAG, (ch + 1) * CW, vb, FB_WCOL(ch + 1));                  \
        }                                                                            \
      }                                                                              \
    } else {                                                                         \
      int ch = sub_t;                                                                \
      if constexpr (FB_EPI_UNROLL == 2) {                                            \
        _Pragma("unroll 1") for (; ch + NW < BN / CW; ch += 2 * NW) {                \
          float v0[CW], v1[CW];                                                      \
          tmem_ld_chunk<CW>(s_addr + ch * CW, v0);                                   \
          tmem_ld_chunk<CW>(s_addr + (ch + NW) * CW, v1);                            \
          tmem_ld_wait();                                                            \
          chunk_body(EDGE_TAG, ch * CW, v0, FB_WCOL(ch));                            \
          chunk_body(EDGE_TAG, (ch + NW) * CW, v1, FB_WCOL(ch + NW));                \
        }                                                                            \
      }                                                                              \
      _Pragma("unroll 1") for (; ch < BN / CW; ch += NW) {                           \
        float v[CW];                                                                 \
        tmem_ld_chunk<CW>(s_addr + ch * CW, v);                                      \
        tmem_ld_wait();                                                              \
        chunk_body(EDGE_TAG, ch * CW, v, FB_WCOL(ch));                               \
      }                                                                              \
    }                                                                                \
  }
        if (diag || tail || padseq || (kIsNce && ptile)) FB_RUN_TILE(std::true_type{})
        else FB_RUN_TILE(std::false_type{})
#undef FB_RUN_TILE
#undef FB_WCOL
        FB_TRACE_EV(2 + wgi, n, 2);
        if (kBwd) {
          tmem_st_wait();
          tc_fence_before();
          mbar_arrive(&bars->w_full[wg]);
        } else {
          tc_fence_before();
          mbar_arrive(&bars->s_empty[wg]);             // S stage drained (all tcgen05.ld of this thread completed)
        }
        if (kColVec) {
          __syncwarp();
          if (lane == 0) mbar_arrive(&bars->b_empty[st]);
        }
      }
      nb += ntiles;
      rowacc = (racc[0] + racc[1]) + (racc[2] + racc[3]);

      if (MODE == NCE_FWD && pos_seen && row_ok)      // exactly one thread of one piece sees the positive of a row
        reinterpret_cast<float*>(ws + p.pos_off)[(((uint64_t)x.q * p.S + x.s) * 2 + side) * p.bpad + row] = posg;
      // ---------------- item epilogue: fold the per-row partials of warpgroups 1.. into warpgroup 0
      if (MODE != NCE_BWD) {
        if (wgi > 0) {
          bars->part_acc[wgi - 1][trow] = rowacc;
          if (!kIsNce) { bars->part_hinge[wgi - 1][trow] = hinge_acc; bars->part_cnt[wgi - 1][trow] = cnt_i; }
        }
        asm volatile("bar.sync 1, %0;" ::"n"(kEpiThreads) : "memory");
        if (wgi == 0) {
#pragma unroll
          for (int w = 0; w < NG - 1; ++w) rowacc += bars->part_acc[w][trow];
          if (MODE == NCE_FWD) {
            const uint64_t slot_stride = (uint64_t)p.nProb * p.S * 2 * p.bpad;
            float* rpart = reinterpret_cast<float*>(ws + p.rpart_off) + (((uint64_t)x.q * p.S + x.s) * 2 + side) * p.bpad;
            rpart[(uint64_t)pk * slot_stride + row] = row_ok ? rowacc : 0.f;
            if (!second)                                                     // slots without a piece read as 0
              for (int k = npi; k < p.nsplit_fwd; ++k) rpart[(uint64_t)k * slot_stride + row] = 0.f;
          } else {
#pragma unroll
            for (int w = 0; w < NG - 1; ++w) { hinge_acc += bars->part_hinge[w][trow]; cnt_i += bars->part_cnt[w][trow]; }
            const bool first_half = !kWide || x.q == 0;         // wide mode: scalars are published by the first pass only
            if (kBwd && row_ok && first_half)
              reinterpret_cast<float*>(ws + p.rho_off + (uint64_t)pk * p.rho2_delta)[(uint64_t)x.c * p.Bpad + row] = rowacc;
            if (row_ok && first_half && (lane & (SQ - 1)) == 0)
              reinterpret_cast<int32_t*>(ws + p.cnt_off + (uint64_t)pk * p.cnt2_delta)[(uint64_t)x.c * p.bpad + seq_i] = cnt_i;
            // every lane of a sequence accumulated the same hinge values: count each (I, J) once
            const float hs = warp_sum(((lane & (SQ - 1)) == 0) ? hinge_acc : 0.f);
            if (lane == 0) bars->red[quarter] = hs;
          }
        }
      }
      if (kBwd) {
        mbar_wait(&bars->o_full, ni & 1);
        tc_fence_after();
        float* out;
        if (kIsNce) {
          out = reinterpret_cast<float*>(ws + p.probs[x.q].dz_off + (uint64_t)pk * p.dz2_delta) +
                ((uint64_t)side * p.S * p.bpad + (uint64_t)x.s * p.bpad + row) * kON;
          if (!second && trow == 0)
            reinterpret_cast<int32_t*>(ws + p.flag_nce_off)[(((uint64_t)x.q * p.S + x.s) * 2 + side) * (p.bpad / kTileM) +
                                                            row0 / kTileM] = npi - 1;
        } else {
          out = reinterpret_cast<float*>(ws + p.dx_off + (uint64_t)pk * p.dx2_delta) +
                ((uint64_t)x.c * p.Bpad + row) * (KB * G::kEPB) + (kWide ? x.q * kON : 0);
          if (!second && trow == 0)
            reinterpret_cast<int32_t*>(ws + p.flag_tmp_off)[(uint64_t)x.c * (p.Bpad / kTileM) + row0 / kTileM] = npi - 1;
        }
        // Each lane holds one row of O (32 columns per tcgen05.ld).  Stored as they are, a warp's 16-byte stores go to 32
        // different 128-byte lines -- 32 L1 tag cycles per instruction, ~12 000 clk per 128 x 256 accumulator
        // (profiles/r2_timeline_r8.txt), paid at the end of every stream-K piece with the tensor pipe idle.  Three
        // lane-bit <-> register-bit exchanges turn the 32 x 32 block so that 8 lanes write the 128 bytes of one row: 4
        // lines per instruction.
        const uint32_t okmask = __ballot_sync(0xffffffffu, row_ok);
        const uint64_t l2_keep = (FB_L2_HINTS & 1) ? l2_policy_evict_last() : 0ull;
        const size_t ostride = kIsNce ? (size_t)kON : (size_t)(KB * G::kEPB);
        float* out_w = out - (size_t)lane * ostride;       // row of lane 0 of this warp
#pragma unroll 1
        for (int ch = wgi; ch < kON / 32; ch += NG) {    // the warpgroups split the columns of O
          float v[32];
          tmem_ld32(tmem + tlane + kOCol + ch * 32, v);
          tmem_ld_wait();
          xchg_lane_reg_bit<0, 2>(v, lane);
          xchg_lane_reg_bit<1, 3>(v, lane);
          xchg_lane_reg_bit<2, 4>(v, lane);
#pragma unroll
          for (int j = 0; j < 8; ++j) {                  // register 4 j + e: row (lane & 24) | j, column 4 (lane & 7) + e
            const int r = (lane & 24) | j;
            if ((okmask >> r) & 1u) {
              float* dst = out_w + (size_t)r * ostride + ch * 32 + 4 * (lane & 7);
              const float4 val = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
              if (FB_L2_HINTS & 1) st4_hint(dst, val, l2_keep);
              else *reinterpret_cast<float4*>(dst) = val;
            }
          }
        }
        tc_fence_before();
        mbar_arrive(&bars->o_empty);
        FB_TRACE_EV(2 + wgi, nb - 1, 3);              // O accumulator of the piece that ended with tile nb - 1 written out
      }
      if (MODE != NCE_BWD) {
        asm volatile("bar.sync 1, %0;" ::"n"(kEpiThreads) : "memory");   // partial arrays / red[] may be reused now
        if (!kIsNce && wgi == 0 && trow == 0 && (!kWide || x.q == 0)) {
          const int t0 = (p.seq0 * p.Sp) / kTileM;
          const int nrt = (p.seq1 * p.Sp + kTileM - 1) / kTileM - t0;
          const int slot = p.np_tmp * (x.c * nrt + (row0 / kTileM - t0));
          float* p3 = reinterpret_cast<float*>(ws + p.part3_off);
          p3[slot + pk] =
              ((bars->red[0] + bars->red[1]) + (bars->red[2] + bars->red[3])) / ((float)p.b * (float)(p.b - 1));
          if (!second)
            for (int k = npi; k < p.np_tmp; ++k) p3[slot + k] = 0.f;
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (threadIdx.x == 0) FB_TRACE_EV(0, 1u, 2);       // kernel exit
  if (warp == 1) tmem_dealloc(tmem, kTmemCols);
}

}  // namespace fb
