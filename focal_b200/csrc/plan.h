// Host-side plan: problem topology (loss.py:162-209 loop structure) and workspace layout in HBM.
#pragma once
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <cstdlib>
#include <cstring>

#include "../../include/focal_b200.h"

#ifdef __CUDACC__
#define FB_HD __host__ __device__ __forceinline__
#else
#define FB_HD inline
#endif

namespace fb {

constexpr int kMaxM = FOCAL_MAX_MODALITIES;
constexpr int kMaxT = 2 * kMaxM;                         // feature tensors: t = view * M + mod
constexpr int kMaxOps = 6 * kMaxM;                       // shared + private (+ full when noPrivate) per tensor
constexpr int kMaxProb = kMaxM * kMaxM;                  // M^2 InfoNCE problems
constexpr int kMaxOrth = 2 * (kMaxM + kMaxM * (kMaxM - 1) / 2);
constexpr int kTileM = 128;                              // rows of a Gram tile (UMMA M, TMEM lanes)
constexpr int kKBlk = 64;                                // bf16 elements per 128-byte swizzle row

struct OpDesc {          // one normalised InfoNCE operand = a column slice of one feature tensor
  int32_t tensor, col0, width, kb;   // kb = K blocks of 128 bytes (64 bf16; split tiles: hi blocks then lo blocks)
  uint64_t off;                      // bytes from ws base: [kb][S*bpad][128 B], position-major rows, swizzled
  int32_t nuse;                      // problems this operand takes part in (<= M - 1 for shared, 1 for private)
  int32_t use_prob[kMaxM];           // ... their indices,
  int32_t use_side[kMaxM];           // ... which side of z the operand is on,
  int32_t use_partner[kMaxM];        // ... and the operand on the other side (row p(k) lives there)
};
struct ProbDesc {        // one InfoNCE problem: z = [opA ; opB] per position (loss.py:66-73)
  int32_t opA, opB, kind;            // kind 0 = shared (modal matching), 1 = private (transformation invariant)
  float weight;
  uint64_t dz_off;                   // fp32 [2][S*bpad][kb*64] operand-gradient accumulators
};
struct OrthDesc {        // one orthogonality pair (loss.py:195-209)
  int32_t tu, cu, tv, cv, width;
};

// Column-tile width of the Gram kernels as a function of the operand width in 64-element K blocks.  TMEM holds
// O (kb*64 columns) + NS S stages (NS*BN): see GramCfg in gram_kernel.cuh.
#ifndef FB_KB4_BN
#define FB_KB4_BN 96
#endif
#ifndef FB_KB8_BN
#define FB_KB8_BN 64
#endif
constexpr int tile_bn(int kb) { return kb <= 2 ? 128 : (kb == 4 ? FB_KB4_BN : (kb == 8 ? FB_KB8_BN : 64)); }

// Shared-memory image of an operand row (128 bytes = 64 bf16 per K block), identical in HBM (the prologue writes it, TMA
// copies it verbatim): SWIZZLE_128B, 16-byte chunk c of row r at c ^ (r & 7).  Byte offset within the row's 128 bytes
// of element `e` of the block.  Split tiles (fp32 mode) store a hi image (blocks [0, kb/2)) and a lo image ([kb/2, kb)).
FB_HD uint32_t tile_byte_bf16(uint32_t row, uint32_t e) { return ((((e >> 3) ^ (row & 7u)) << 4) | ((e & 7u) << 1)); }

// Problems handled by one InfoNCE launch (all with the same operand width / K-block count).
struct ProbSel {
  int32_t n;
  int32_t idx[kMaxProb];
  int32_t peer_wait;   // row-sharded path: number of ranks whose barrier announcement this launch waits for (0 = none)
  // row-sharded path: block 0 first announces the rank's epoch (the stores of the previous launch are complete) to the
  // ann_world mapped workspaces in peer_ws (0 = nothing to announce)
  int32_t ann_world, ann_rank;
  uint8_t* peer_ws[8];
};

struct Plan {
  // dims
  int32_t B, S, M, D, d, b, bpad, Bpad, nT, nOps, nProb, nOrth, kbFull, seq0, seq1, need_grad, terms, num_sms;
  // Temporal row space: the distance kernel reduces S x S blocks with shuffles over Sp = the next power of two >= S lanes,
  // so every temporal per-row array (operands, squared norms, m_II, dx, rho) is indexed by tr(i) = (i / S) * Sp + i % S;
  // rows with position >= S are phantoms (zero operands, masked everywhere).  Sp == S for the usual power-of-two S.
  int32_t Sp, Bt;                                       // padded sequence length, rows of the temporal row space (b * Sp)
  int32_t nsplit_fwd;                                   // row-sum slots per row (= np_nce)
  // stream-K split of the Gram launches (see PieceIter): grid sizes and the largest number of pieces one 128-row
  // block is cut into; piece k of a block accumulates into the k-th copy of the output buffers
  int32_t sk_nce, sk_tmp, np_nce, np_tmp, grid_tmp, grid_nce[9];
  int32_t in_rb, in_bs;                                 // row-blocked inputs: rows per block, block stride (floats)
  int32_t local_rows;                                   // sharded path: the prologue handles the owned rows only
  int32_t prec, epb;                                    // tile precision (FOCAL_PREC_*), operand COLUMNS per K block (64 / 32)
  int32_t wide;                                         // bf16, 256 < D <= 512: O accumulator in two 256-column passes
  int32_t indirect;                                     // caller pointers come from the table at ptrs_off (CUDA graphs)
  // Row-kernel generation: 3 = row_kernels_v3.cuh (a warp per (sequence, tensor); seqb sequences x nT tensors per block,
  // nq float4 slots per lane and half), 0 = the older kernels (focal_b200.cu picks among them)
  int32_t rowgen, seqb, nq;
  float T, margin, w_shared, w_private, w_orth, w_rank;
  float alpha;                                          // sqrt(log2(e)/T): operand pre-scale, Gram = log2-domain logit
  OpDesc ops[kMaxOps];
  ProbDesc probs[kMaxProb];
  OrthDesc orth[kMaxOrth];
  // workspace offsets (bytes)
  uint64_t xt_off;      // bf16 [2M][kbFull][Bpad][64] temporal operands (natural row order, swizzled)
  uint64_t sq_off;      // fp32 [2M][Bpad] squared norms of the rounded temporal operands
  uint64_t mintra_off;  // fp32 [2M][Bpad] m_II of the row's sequence (exact fp32)
  uint64_t nrm_off;     // fp32 [2M][Bpad][2] squared norms of the shared / private half (prologue_v2 -> finalize_v2)
  uint64_t pd_off;      // fp32 [2M][Bpad][4] squared distances of the rounded row to rows g^1, g^2, g^3 of its sequence (v3)
  uint64_t rpart_off;   // fp32 [nsplit_fwd][nProb][S][2][bpad]
  uint64_t rsum_off;    // fp32 [nProb][S][2][bpad]
  uint64_t rinv_off;    // fp32 [nProb][S][2][bpad]
  uint64_t pos_off;     // fp32 [nProb][S][2][bpad] positive-pair logit G_{k,p(k)} (log2 domain) as the row-sum tile saw it
  uint64_t dx_off;      // fp32 [2M][Bpad][kbFull*epb]
  uint64_t rho_off;     // fp32 [2M][Bpad] sum_j r_ij
  uint64_t cnt_off;     // int32 [2M][bpad]  active hinges per sequence
  uint64_t part1_off;   // fp32 [nblk1][4]  prologue partials: orth
  uint64_t part2_off;   // fp32 [nblk2][2]  log-row-sum partials: shared, private
  uint64_t part3_off;   // fp32 [nitems3]   hinge partials
  uint64_t lossd_off;   // double [8]
  uint64_t dz_off, dz_bytes, dx_bytes;
  // stream-K: a 128-row block whose column tiles are split over several CTAs has one set of accumulators per piece
  uint64_t dz2_delta, dx2_delta, rho2_delta, cnt2_delta;   // byte distance between the buffers of consecutive pieces
  uint64_t flag_tmp_off;   // int32 [2M][Bpad/128]: number of secondary pieces of the row block (temporal)
  uint64_t flag_nce_off;   // int32 [nProb][S][2][bpad/128]: same for the InfoNCE backward pass
  uint64_t bar_off;        // uint32 [16]: [r] = last barrier epoch rank r announced here, [8] = own epoch counter,
                           //              [9] = blocks-done counter of the launch that announces next
  uint64_t lossx_off;      // double [kMaxPeers][8]: loss partials of every rank (sharded path)
  uint64_t ptrs_off;       // PtrTable: the caller's feature / gradient / loss pointers of this step (indirect mode)
  uint64_t total_bytes;
  int32_t nblk1, nblk2, nitems3;
};

// Caller-owned pointers of one step.  Plain launches pass them by value as kernel arguments; in indirect mode
// (FocalCfg.indirect_ptrs, used when the launch sequence is replayed from a CUDA graph) the row kernels read them from
// this table in the workspace, which focal_b200_set_ptrs refreshes before every replay -- so a captured graph serves
// freshly allocated inputs and outputs every step.
struct PtrTable {
  const float* x[kMaxT];
  float* g[kMaxT];
  float* loss5;
};

// Workspaces of all ranks of a row-sharded job as mapped into this process (plain launches: world = 1, ws[0] = own).
constexpr int kMaxPeers = 8;
struct PeerWs {
  int32_t rank, world;
  uint8_t* ws[kMaxPeers];
  uint8_t* mc;                // NVSwitch multicast mapping of all workspaces (nullptr: store to every ws[r] instead)
};

inline uint64_t align_up(uint64_t x, uint64_t a) { return (x + a - 1) / a * a; }

// element offset of row i inside a (possibly row-blocked) feature tensor
FB_HD size_t feat_row_off(const Plan& p, int i) {
  return (size_t)(i / p.in_rb) * (size_t)p.in_bs + (size_t)(i % p.in_rb) * (size_t)p.D;
}

// rows of the prologue / finalize kernels handled per 128-thread block (one warp per row)
constexpr int kRowsPerBlock = 4;

// Stream-K geometry of one Gram launch: n_items row blocks of T column tiles each are numbered consecutively and cut
// into one contiguous share per CTA.  Returns the grid size and (np) the largest number of pieces a row block is cut
// into; the grid shrinks until np <= kMaxPieces and a share is at least kMinShare tiles.
constexpr int kMaxPieces = 8;
constexpr int kMinShare = 4;
inline int piece_grid(int n_items, int T, int num_sms, bool streamk, int* np) {
  *np = 1;
  if (n_items <= 0) return 0;
  if (!streamk) return n_items < num_sms ? n_items : num_sms;
  const long total = (long)n_items * T;
  long grid = num_sms;
  if (total / kMinShare < grid) grid = total / kMinShare > 0 ? total / kMinShare : 1;
  for (;; --grid) {
    const long share = (total + grid - 1) / grid;
    int worst = 1;
    for (int i = 0; i < n_items; ++i) {
      const long i0 = (long)i * T;
      const int n = (int)((i0 + T - 1) / share - i0 / share) + 1;
      if (n > worst) worst = n;
    }
    if (worst <= kMaxPieces || grid == 1) { *np = worst; return (int)grid; }
  }
}
// The share of one CTA of a stream-K launch, piece by piece (device: the three roles of gram_kernel walk it in step;
// host: tests/csrc/test_plan.cpp checks that the pieces of all CTAs tile the launch exactly once).
struct PieceIter {
  long u, u1, share;
  int T, item_, n_items_, cta_, grid_;
  bool streamk;
  FB_HD PieceIter(int cta, int grid, int n_items, int tiles_per_item, bool use_streamk) {
    T = tiles_per_item;
    streamk = use_streamk;
    n_items_ = n_items;
    item_ = cta; cta_ = cta; grid_ = grid;
    const long total = (long)n_items * T;
    share = (total + grid - 1) / grid;
    u = (long)cta * share;
    u1 = u + share < total ? u + share : total;
  }
  // next piece of this CTA: row block `item`, column tiles [t0, t1); pk = index of the piece within its row block,
  // npi = number of pieces the row block is cut into
  FB_HD bool next(int& item, int& t0, int& t1, int& pk, int& npi) {
    if (!streamk) {                       // whole row blocks, CTA-strided
      if (item_ >= n_items_) return false;
      item = item_; t0 = 0; t1 = T; pk = 0; npi = 1;
      item_ += grid_;
      return true;
    }
    if (u >= u1) return false;
    item = (int)(u / T);
    const long i0 = (long)item * T;
    t0 = (int)(u - i0);
    const long rest = u1 - u;
    t1 = (long)(T - t0) <= rest ? T : (int)(t0 + rest);
    const int cfirst = (int)(i0 / share);
    pk = cta_ - cfirst;
    npi = (int)((i0 + T - 1) / share) - cfirst + 1;
    u += t1 - t0;
    return true;
  }
};

FB_HD int nce_row_tiles(const Plan& p) { return (p.seq1 + kTileM - 1) / kTileM - p.seq0 / kTileM; }
FB_HD int tmp_row_tiles(const Plan& p) {
  return (p.seq1 * p.Sp + kTileM - 1) / kTileM - (p.seq0 * p.Sp) / kTileM;
}
// row of the temporal row space that holds feature row i
FB_HD int tmp_row(const Plan& p, int i) { return p.Sp == p.S ? i : (i / p.S) * p.Sp + i % p.S; }
#ifndef FB_ROW_V3
#define FB_ROW_V3 1             // 0: experiment builds without the third-generation row kernels (A/B measurements)
#endif
#ifndef FB_STREAMK_TMP
#define FB_STREAMK_TMP 1        // 1: always; 0: only when the launch has fewer row blocks than SMs (row shards); -1: never
#endif
#ifndef FB_STREAMK_NCE
#define FB_STREAMK_NCE 1
#endif

// FOCAL_B200_ROW_KERNELS=v1|v2 in the environment: keep the older row kernels (A/B measurements)
inline bool row_kernels_forced_old() {
  const char* e = std::getenv("FOCAL_B200_ROW_KERNELS");
  return e && (std::strcmp(e, "v1") == 0 || std::strcmp(e, "v2") == 0);
}

inline int build_plan(const FocalCfg& c, Plan& p, int num_sms) {
  std::memset(&p, 0, sizeof(p));
  if (c.B <= 0 || c.S <= 0 || c.M <= 0 || c.D <= 0) return FOCAL_EINVAL;
  if (c.M < 1 || c.M > kMaxM) return FOCAL_ESHAPE;
  if (c.B % c.S) return FOCAL_ESHAPE;                    // loss.py:154 reshape(-1, S, D) would raise
  if (c.S > 32) return FOCAL_ESHAPE;                     // a (padded) sequence = rows of one warp
  if (c.D < 2 || c.D > 512) return FOCAL_ESHAPE;
  if (!(c.temperature > 0.f)) return FOCAL_EINVAL;
  if (c.precision != FOCAL_PREC_BF16 && c.precision != FOCAL_PREC_FP32) return FOCAL_EINVAL;
  p.prec = c.precision;
  p.epb = c.precision == FOCAL_PREC_FP32 ? 32 : kKBlk;
  // split tiles are 4 bytes per element: a 128-row A tile of more than 256 columns does not fit shared memory
  if (c.precision == FOCAL_PREC_FP32 && c.D > 256) return FOCAL_ESHAPE;
  p.B = c.B; p.S = c.S; p.M = c.M; p.D = c.D; p.d = c.D / 2;
  p.b = c.B / c.S;
  p.Sp = 1;
  while (p.Sp < c.S) p.Sp *= 2;
  p.Bt = p.b * p.Sp;
  p.nT = 2 * c.M;
  // K blocks (128 bytes = 64 bf16 elements, zero-padded) of an operand of `w` columns: bf16 1..4, or the wide mode's 8;
  // split tiles: a hi and a lo image of ceil(w / 64) blocks each -> 2, 4, 6, 8
  auto kblocks = [&](int w) {
    const int kb = (w + kKBlk - 1) / kKBlk;
    if (p.prec == FOCAL_PREC_BF16) return kb > 4 ? 8 : kb;
    return 2 * kb;
  };
  // temporal operand: bf16 1..4 blocks, or 8 (zero-padded) for 256 < D <= 512, the "wide" Gram mode; split 2..8
  p.kbFull = kblocks(c.D);
  p.wide = (c.precision == FOCAL_PREC_BF16 && p.kbFull > 4) ? 1 : 0;
  {
    // rows are padded so that every 128-row A tile and every BN-row B tile stays inside the operand arrays
    auto pad_rows = [](int rows, int bn) {
      const int by_tile = (rows + bn - 1) / bn * bn;
      return (int32_t)align_up((uint64_t)(by_tile > rows ? by_tile : rows), kTileM);
    };
    const int kbHalf = kblocks(p.d);
    p.bpad = pad_rows(p.b, tile_bn(kbHalf));
    if (c.no_private) { const int32_t alt = pad_rows(p.b, tile_bn(p.kbFull)); if (alt > p.bpad) p.bpad = alt; }
    p.Bpad = pad_rows(p.Bt, tile_bn(p.kbFull));
  }
  p.seq0 = c.seq_begin; p.seq1 = c.seq_end;
  if (p.seq0 < 0 || p.seq1 > p.b || p.seq0 >= p.seq1) return FOCAL_EINVAL;
  p.need_grad = c.need_grad; p.terms = c.terms ? c.terms : FOCAL_TERM_ALL;
  p.in_rb = c.in_block_rows > 0 ? c.in_block_rows : c.B;
  p.in_bs = c.in_block_rows > 0 ? c.in_block_stride : 0;
  if (p.in_rb % c.S || c.B % p.in_rb) return FOCAL_EINVAL;
  p.local_rows = c.local_rows ? 1 : 0;
  p.indirect = c.indirect_ptrs ? 1 : 0;
  if (p.local_rows && c.in_block_rows > 0) return FOCAL_EINVAL;
  p.num_sms = num_sms;
  p.T = c.temperature; p.margin = c.margin;
  p.w_shared = c.w_shared; p.w_private = c.w_private; p.w_orth = c.w_orth; p.w_rank = c.w_rank;
  p.alpha = sqrtf(1.4426950408889634f / c.temperature);
  // log2-domain logits are bounded by log2(e)/T; without a running max the row sum must fit fp32
  if (1.4426950408889634f / c.temperature > 96.f) return FOCAL_ESHAPE;

  // ---- operands: per tensor [shared, private] (+ full when noPrivate)
  const int M = c.M, d = p.d;
  auto op_shared = [&](int t) { return 2 * t; };
  auto op_private = [&](int t) { return 2 * t + 1; };
  auto op_full = [&](int t) { return 2 * p.nT + t; };
  for (int t = 0; t < p.nT; ++t) {
    p.ops[op_shared(t)] = OpDesc{t, 0, d, kblocks(d), 0, 0, {0}, {0}, {0}};
    p.ops[op_private(t)] = OpDesc{t, d, d, kblocks(d), 0, 0, {0}, {0}, {0}};
  }
  p.nOps = 2 * p.nT;
  if (c.no_private) {
    for (int t = 0; t < p.nT; ++t) p.ops[op_full(t)] = OpDesc{t, 0, c.D, p.kbFull, 0, 0, {0}, {0}, {0}};
    p.nOps = 3 * p.nT;
  }
  // ---- problems in reference order (loss.py:162-186)
  int np = 0;
  for (int v = 0; v < 2; ++v)
    for (int i = 0; i < M; ++i)
      for (int j = i + 1; j < M; ++j) {
        int ta = v * M + i, tb = v * M + j;
        p.probs[np++] = ProbDesc{c.no_private ? op_full(ta) : op_shared(ta), c.no_private ? op_full(tb) : op_shared(tb),
                                 0, c.w_shared, 0};
      }
  for (int m = 0; m < M; ++m) p.probs[np++] = ProbDesc{op_private(m), op_private(M + m), 1, c.w_private, 0};
  p.nProb = np;
  // ---- orthogonality pairs (loss.py:195-209)
  int no = 0;
  for (int v = 0; v < 2; ++v)
    for (int i = 0; i < M; ++i) {
      int t = v * M + i;
      p.orth[no++] = OrthDesc{t, 0, t, d, d};
      for (int j = i + 1; j < M; ++j) p.orth[no++] = OrthDesc{t, d, v * M + j, d, d};
    }
  p.nOrth = no;
  // per-operand use lists (what finalize needs: which dz accumulators belong to an operand)
  for (int o = 0; o < p.nOps; ++o) p.ops[o].nuse = 0;
  for (int q = 0; q < p.nProb; ++q) {
    const int oa = p.probs[q].opA, ob = p.probs[q].opB;
    OpDesc& A = p.ops[oa];
    OpDesc& Bo = p.ops[ob];
    A.use_prob[A.nuse] = q; A.use_side[A.nuse] = 0; A.use_partner[A.nuse] = ob; ++A.nuse;
    Bo.use_prob[Bo.nuse] = q; Bo.use_side[Bo.nuse] = 1; Bo.use_partner[Bo.nuse] = oa; ++Bo.nuse;
  }

  // ---- stream-K geometry of the Gram launches
  p.sk_nce = FB_STREAMK_NCE;
  p.np_nce = 2;
  for (int kb = 1; kb <= 8; ++kb) {
    int n = 0, np1 = 1;
    for (int q = 0; q < p.nProb; ++q) n += p.ops[p.probs[q].opA].kb == kb;
    if (!n) continue;
    if (kb > 4) return FOCAL_ESHAPE;         // InfoNCE operands wider than 4 K blocks (noPrivate: D > 256 bf16, > 128 split)
    const int bn = tile_bn(kb);
    p.grid_nce[kb] = piece_grid(n * p.S * 2 * nce_row_tiles(p), 2 * ((p.b + bn - 1) / bn), num_sms, p.sk_nce != 0, &np1);
    if (np1 > p.np_nce) p.np_nce = np1;
  }
  {
    // wide mode (kbFull == 8): the O accumulator holds half of the columns, so the backward launch visits every row
    // block twice (one item per output half) and keeps whole row blocks per CTA (no stream-K pieces)
    const bool wide = p.wide != 0;
    const int items = p.nT * tmp_row_tiles(p) * ((wide && p.need_grad) ? 2 : 1), bn = tile_bn(p.kbFull);
    p.sk_tmp = !wide && ((FB_STREAMK_TMP > 0) || (FB_STREAMK_TMP == 0 && items < num_sms));
    p.np_tmp = 2;
    int np1 = 1;
    p.grid_tmp = piece_grid(items, (p.Bt + bn - 1) / bn, num_sms, p.sk_tmp != 0, &np1);
    if (np1 > p.np_tmp) p.np_tmp = np1;
  }
  p.nsplit_fwd = p.np_nce;      // row-sum slots per row: one per stream-K piece

  // ---- workspace
  uint64_t off = 0;
  auto take = [&](uint64_t bytes) { uint64_t o = off; off = align_up(off + bytes, 1024); return o; };
  const uint64_t rowsNce = (uint64_t)p.S * p.bpad;
  for (int o = 0; o < p.nOps; ++o) p.ops[o].off = take((uint64_t)p.ops[o].kb * rowsNce * 128);
  p.xt_off = take((uint64_t)p.nT * p.kbFull * p.Bpad * 128);
  p.sq_off = take((uint64_t)p.nT * p.Bpad * 4);
  p.mintra_off = take((uint64_t)p.nT * p.Bpad * 4);
  p.nrm_off = take((uint64_t)p.nT * p.Bpad * 8);
  p.pd_off = take((uint64_t)p.nT * p.Bpad * 16);
  const uint64_t rs = (uint64_t)p.nProb * p.S * 2 * p.bpad * 4;
  p.rpart_off = take(rs * p.nsplit_fwd);
  p.rsum_off = take(rs);
  p.rinv_off = take(rs);
  p.pos_off = take(rs);
  p.dz_off = off;
  for (int q = 0; q < p.nProb; ++q)
    p.probs[q].dz_off = take((uint64_t)2 * rowsNce * p.ops[p.probs[q].opA].kb * p.epb * 4);
  p.dz_bytes = off - p.dz_off;
  p.dz2_delta = p.dz_bytes;
  for (int k = 1; k < p.np_nce; ++k) take(p.dz_bytes);               // dz accumulators of the secondary pieces
  auto take_n = [&](uint64_t bytes, int n, uint64_t* delta) {        // n equally spaced copies
    const uint64_t o = take(bytes);
    *delta = align_up(bytes, 1024);
    for (int k = 1; k < n; ++k) take(bytes);
    return o;
  };
  p.dx_bytes = (uint64_t)p.nT * p.Bpad * p.kbFull * p.epb * 4;
  p.dx_off = take_n(p.dx_bytes, p.np_tmp, &p.dx2_delta);
  p.rho_off = take_n((uint64_t)p.nT * p.Bpad * 4, p.np_tmp, &p.rho2_delta);
  p.cnt_off = take_n((uint64_t)p.nT * p.bpad * 4, p.np_tmp, &p.cnt2_delta);
  p.flag_tmp_off = take((uint64_t)p.nT * (p.Bpad / kTileM) * 4);
  p.flag_nce_off = take((uint64_t)p.nProb * p.S * 2 * (p.bpad / kTileM) * 4);
  p.nblk1 = ((p.local_rows ? (p.seq1 - p.seq0) * p.S : p.B) + kRowsPerBlock - 1) / kRowsPerBlock;
  // third-generation row kernels: standard topology, S in {1, 2, 4}, a whole number (<= 4) of float4 slots per lane and
  // half (nq = d * S / 128), D <= 256, up to 8 tensors; their blocks cover seqb sequences each
  p.rowgen = 0; p.seqb = 0; p.nq = 0;
  if (!c.no_private && (c.D & 1) == 0 && (c.S == 1 || c.S == 2 || c.S == 4) && (p.d * c.S) % 128 == 0 &&
      p.d * c.S / 128 <= 4 && c.D <= 256 && p.nT <= 8 && FB_ROW_V3 && !row_kernels_forced_old()) {
    p.rowgen = 3;
    p.nq = p.d * c.S / 128;
    p.seqb = 8 / p.nT;
    p.nblk1 = ((p.local_rows ? (p.seq1 - p.seq0) : p.b) + p.seqb - 1) / p.seqb;
  }
  p.nblk2 = (int32_t)((rowsNce * 2 * p.nProb + 255) / 256);
  p.nitems3 = p.np_tmp * p.nT * (p.Bpad / kTileM);   // one hinge-partial slot per piece of a row block
  p.part1_off = take((uint64_t)p.nblk1 * 4 * 4);
  p.part2_off = take((uint64_t)p.nblk2 * 2 * 4);
  p.part3_off = take((uint64_t)p.nitems3 * 4);
  p.lossd_off = take(8 * 8);
  p.bar_off = take(16 * 4);
  p.lossx_off = take((uint64_t)kMaxPeers * 8 * 8);
  p.ptrs_off = take(sizeof(PtrTable));
  p.total_bytes = off;
  return FOCAL_OK;
}

}  // namespace fb
