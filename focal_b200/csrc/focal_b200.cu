// C ABI of the B200-native FOCAL loss hot path (see include/focal_b200.h).
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <unordered_map>
#include <cuda_runtime.h>

#include "../../include/focal_b200.h"
#include "augment_kernels.cuh"
#include "gram_kernel.cuh"
#include "knn_kernels.cuh"
#include "plan.h"
#include "row_kernels.cuh"
#include "row_kernels_fast.cuh"
#include "row_kernels_v2.cuh"
#include "row_kernels_v3.cuh"

using namespace fb;

namespace {

int cuda_ok(const char* what);

int device_sms() {
  static int cached[64] = {0};          // SM count per device ordinal (never changes; avoid a driver query per call)
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (cached[dev] <= 0) {
    int sms = 0;
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) return 148;
    cached[dev] = sms;
  }
  return cached[dev];
}

// Every entry point rebuilds the plan from the cfg; the last few are cached per thread (the stream-K geometry search in
// build_plan is the expensive part, and a step calls six entry points with the same cfg).
int make_plan(const FocalCfg* cfg, Plan& p) {
  if (!cfg) return FOCAL_EINVAL;
  const int sms = cfg->num_sms > 0 ? cfg->num_sms : device_sms();
  struct Slot { FocalCfg cfg; int sms; int rc; bool valid; Plan plan; };
  constexpr int kSlots = 8;
  thread_local Slot slots[kSlots];
  thread_local int next = 0;
  for (int i = 0; i < kSlots; ++i)
    if (slots[i].valid && slots[i].sms == sms && std::memcmp(&slots[i].cfg, cfg, sizeof(FocalCfg)) == 0) {
      if (slots[i].rc == FOCAL_OK) p = slots[i].plan;
      return slots[i].rc;
    }
  Slot& s = slots[next];
  next = (next + 1) % kSlots;
  s.cfg = *cfg; s.sms = sms; s.valid = true;
  s.rc = build_plan(*cfg, s.plan, sms);
  if (s.rc == FOCAL_OK) p = s.plan;
  return s.rc;
}

// cudaFuncAttributeMaxDynamicSharedMemorySize is a per-device (per-context) attribute of ONE kernel: remember what was
// set per (kernel, device ordinal), so a second GPU driven from the same process gets its own call.  (Keyed on the
// function address: all instantiations of a kernel template share one function-pointer TYPE.)
template <class K>
int ensure_dyn_smem(K kfn, size_t bytes, const char* what) {
  if (bytes <= 48 * 1024) return FOCAL_OK;
  static std::mutex mu;
  static std::unordered_map<uint64_t, size_t> configured;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0) dev = 0;
  const uint64_t key = (uint64_t)reinterpret_cast<uintptr_t>(reinterpret_cast<const void*>(kfn)) * 64u + (uint64_t)(dev & 63);
  std::lock_guard<std::mutex> lock(mu);
  size_t& have = configured[key];
  if (bytes > have) {
    if (cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes) != cudaSuccess) return cuda_ok(what);
    have = bytes;
  }
  return FOCAL_OK;
}

int check_ws(const Plan& p, const void* ws, size_t ws_bytes) {
  if (!ws) return FOCAL_EINVAL;
  if (reinterpret_cast<uintptr_t>(ws) & 1023) return FOCAL_EWORKSPACE;
  if (ws_bytes < p.total_bytes) return FOCAL_EWORKSPACE;
  return FOCAL_OK;
}

int cuda_ok(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    fprintf(stderr, "focal_b200: %s failed: %s\n", what, cudaGetErrorString(e));
    return FOCAL_ECUDA;
  }
  return FOCAL_OK;
}

bool temporal_degenerate(const Plan& p) { return p.b <= 1 || p.S <= 1; }

// blocks of nce_lse_kernel: the owned (problem, position, side, sequence) entries, or every entry when all_rows
int lse_blocks(const Plan& p, int all_rows) {
  if (all_rows) return p.nblk2;
  return (int)(((long)p.nProb * p.S * 2 * (p.seq1 - p.seq0) + 255) / 256);
}

// Launch with programmatic stream serialisation: the kernel may be scheduled while the launch before it on the stream
// drains.  ONLY for kernels that call pdl_wait() before their first global-memory access (gram_kernel, prologue_v3,
// finalize_v3, nce_lse).  OFF unless FOCAL_B200_PDL=1 is in the environment: measured (profiles/r2_pdl_ab.txt) it makes
// the step SLOWER -- headline 515.7 -> 523.8 us, cfg2 72.7 -> 80.1 us, cfg1 80.9 -> 82.9 us: the early-resident blocks
// of the next launch only take SM slots from the tail of the current one, and there is no set-up worth overlapping.
template <class... KArgs, class... Args>
cudaError_t launch_pdl(void (*kfn)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  static const bool on = [] { const char* e = std::getenv("FOCAL_B200_PDL"); return e && e[0] == '1'; }();
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = on ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kfn, static_cast<KArgs>(args)...);
}

template <int MODE, int KB, int SEQ, int EL>
int launch_gram(const Plan& p, const ProbSel& sel, uint8_t* ws, cudaStream_t st, int grid) {
  using G = GramCfg<MODE, KB, SEQ, EL>;
  static_assert(G::BN == tile_bn(KB), "plan.h and gram_kernel.cuh must agree on the column tile");
  using L = GramSmem<G::BN, KB, G::NB>;
  auto kfn = gram_kernel<MODE, KB, SEQ, EL>;
  if (int rc = ensure_dyn_smem(kfn, L::kDynamic, "cudaFuncSetAttribute(gram_kernel)")) return rc;
  if (grid <= 0) return FOCAL_OK;                                 // persistent: at most one CTA per SM (plan.h)
  launch_pdl(kfn, dim3(grid), dim3(G::kThreads), L::kDynamic, st, p, sel, ws);
  return cuda_ok("gram_kernel launch");
}

// operand width in 128-byte K blocks.  bf16 tiles: 1..4, temporal launches also 8 ("wide": 256 < D <= 512).
// Split tiles (hi + lo images): 2, 4, temporal launches also 6 (D = 192) and 8 (D = 256).
template <int MODE, int SEQ, int EL>
int launch_gram_kb(const Plan& p, const ProbSel& sel, uint8_t* ws, cudaStream_t st, int kb, int grid) {
  switch (kb) {
#ifndef FB_FAST_BUILD
    case 1:
      if constexpr (EL == 0) return launch_gram<MODE, 1, SEQ, EL>(p, sel, ws, st, grid);
      break;
    case 3:
      if constexpr (EL == 0) return launch_gram<MODE, 3, SEQ, EL>(p, sel, ws, st, grid);
      break;
    case 6:
      if constexpr (EL == 1 && MODE >= 2) return launch_gram<MODE, 6, SEQ, EL>(p, sel, ws, st, grid);
      break;
#endif
    case 2: return launch_gram<MODE, 2, SEQ, EL>(p, sel, ws, st, grid);
    case 4: return launch_gram<MODE, 4, SEQ, EL>(p, sel, ws, st, grid);
    case 8:
#ifdef FB_FAST_BUILD
      if constexpr (EL == 1 && MODE >= 2) return launch_gram<MODE, 8, SEQ, EL>(p, sel, ws, st, grid);
#else
      if constexpr (MODE >= 2) return launch_gram<MODE, 8, SEQ, EL>(p, sel, ws, st, grid);
#endif
      break;
  }
  return FOCAL_ESHAPE;
}
template <int MODE, int SEQ>
int launch_gram_prec(const Plan& p, const ProbSel& sel, uint8_t* ws, cudaStream_t st, int kb, int grid) {
  return p.prec == FOCAL_PREC_FP32 ? launch_gram_kb<MODE, SEQ, 1>(p, sel, ws, st, kb, grid)
                                   : launch_gram_kb<MODE, SEQ, 0>(p, sel, ws, st, kb, grid);
}

template <int MODE>
int launch_temporal(const Plan& p, uint8_t* ws, cudaStream_t st, int grid, const ProbSel& sync = ProbSel{}) {
  ProbSel sel = sync;                 // peer announce / wait of the row-sharded path (none otherwise)
  switch (p.Sp) {                     // lanes per (padded) sequence; p.S itself may be any length up to 32
    case 4: return launch_gram_prec<MODE, 4>(p, sel, ws, st, p.kbFull, grid);
#ifndef FB_FAST_BUILD                 // experiment builds (tools/variant_bench.py) only instantiate the headline shapes
    case 2: return launch_gram_prec<MODE, 2>(p, sel, ws, st, p.kbFull, grid);
    case 8: return launch_gram_prec<MODE, 8>(p, sel, ws, st, p.kbFull, grid);
    case 16: return launch_gram_prec<MODE, 16>(p, sel, ws, st, p.kbFull, grid);
    case 32: return launch_gram_prec<MODE, 32>(p, sel, ws, st, p.kbFull, grid);
#endif
  }
  return FOCAL_ESHAPE;
}

// InfoNCE problems are launched in groups of equal operand width (they differ only under noPrivate, where the
// shared problems use the full D columns and the private ones D/2).
template <int MODE>
int launch_nce(const Plan& p, uint8_t* ws, cudaStream_t st, const ProbSel& sync = ProbSel{}) {
  bool first = true;
  for (int kb = 1; kb <= 4; ++kb) {
    ProbSel sel = sync;               // peer announce / wait of the row-sharded path (none otherwise)
    sel.n = 0;
    if (!first) sel.ann_world = 0;    // one announcement per producer launch
    for (int q = 0; q < p.nProb; ++q)
      if (p.ops[p.probs[q].opA].kb == kb) sel.idx[sel.n++] = q;
    if (!sel.n) continue;
    const int rc = launch_gram_prec<MODE, 0>(p, sel, ws, st, kb, p.grid_nce[kb]);
    if (rc) return rc;
    first = false;
  }
  return FOCAL_OK;
}

// lanes own VW = d/32 consecutive columns per half on the vectorised row-kernel path (0 = use the generic kernels)
int fast_row_vw(const Plan& p, int no_private) {
  if (no_private || (p.D & 1) || p.d % 32 || p.d < 32 || p.Sp != p.S) return 0;
  const int vw = p.d / 32;
  return (vw <= 4 || vw == 8) ? vw : 0;           // D = 64, 128, 192, 256, 512
}

template <int VW, int PREC>
int launch_prologue_fast_vw(const Plan& p, const FeatPtrs& f, const PeerWs& pw, uint8_t* w, size_t smem, int grid, int fuse,
                            cudaStream_t st) {
  if (int rc = ensure_dyn_smem(prologue_fast_kernel<VW, PREC>, smem, "cudaFuncSetAttribute(prologue_fast_kernel)")) return rc;
  prologue_fast_kernel<VW, PREC><<<grid, 128, smem, st>>>(p, f, pw, w, fuse);
  return cuda_ok("prologue_fast_kernel");
}
template <int PREC>
int launch_prologue_fast_p(int vw, const Plan& p, const FeatPtrs& f, const PeerWs& pw, uint8_t* w, size_t smem, int grid,
                           int fuse, cudaStream_t st) {
  switch (vw) {
    case 1: return launch_prologue_fast_vw<1, PREC>(p, f, pw, w, smem, grid, fuse, st);
    case 2: return launch_prologue_fast_vw<2, PREC>(p, f, pw, w, smem, grid, fuse, st);
    case 3: return launch_prologue_fast_vw<3, PREC>(p, f, pw, w, smem, grid, fuse, st);
    case 4: return launch_prologue_fast_vw<4, PREC>(p, f, pw, w, smem, grid, fuse, st);
    case 8:
      if constexpr (PREC == FOCAL_PREC_BF16) return launch_prologue_fast_vw<8, PREC>(p, f, pw, w, smem, grid, fuse, st);
      break;                                      // split tiles stop at D = 256
  }
  return FOCAL_ESHAPE;
}
int launch_prologue_fast(int vw, const Plan& p, const FeatPtrs& f, const PeerWs& pw, uint8_t* w, size_t smem, int grid,
                         int fuse, cudaStream_t st) {
  return p.prec == FOCAL_PREC_FP32 ? launch_prologue_fast_p<FOCAL_PREC_FP32>(vw, p, f, pw, w, smem, grid, fuse, st)
                                   : launch_prologue_fast_p<FOCAL_PREC_BF16>(vw, p, f, pw, w, smem, grid, fuse, st);
}
#ifndef FB_FINALIZE_RT
#define FB_FINALIZE_RT 1          // 1: one warp per (row, tensor) when the launch is latency-bound; 0: always one warp per row
#endif
template <int VW, int MAXT, int PREC>
int launch_finalize_rt_t(const Plan& p, const FeatPtrs& f, const GradPtrs& g, const uint8_t* w, int grid, cudaStream_t st) {
  const size_t smem = ((size_t)4 * p.nT * p.D + 4 * 2 * kMaxT) * sizeof(float);
  if (int rc = ensure_dyn_smem(finalize_rt_kernel<VW, MAXT, PREC>, smem, "cudaFuncSetAttribute(finalize_rt_kernel)")) return rc;
  finalize_rt_kernel<VW, MAXT, PREC><<<grid, MAXT > 0 ? 128 * p.nT : 128, smem, st>>>(p, f, g, w);
  return cuda_ok("finalize_rt_kernel");
}
template <int VW, int PREC>
int launch_finalize_vw(const Plan& p, const FeatPtrs& f, const GradPtrs& g, const uint8_t* w, int grid, cudaStream_t st,
                       bool rt) {
  if (!rt) return launch_finalize_rt_t<VW, 0, PREC>(p, f, g, w, grid, st);
  // <= 4 tensors (M <= 2): 512-thread blocks, 128 registers per thread available; else 1024-thread blocks
  return p.nT <= 4 ? launch_finalize_rt_t<VW, 4, PREC>(p, f, g, w, grid, st)
                   : launch_finalize_rt_t<VW, 8, PREC>(p, f, g, w, grid, st);
}
template <int PREC>
int launch_finalize_fast_p(int vw, const Plan& p, const FeatPtrs& f, const GradPtrs& g, const uint8_t* w, int grid,
                           cudaStream_t st) {
  // Few rows (a row shard): the launch is one wave of blocks and its time is the dependent chain of one warp -> one warp
  // per (row, tensor) (measured 1024 rows: 41 -> 32 us).  Many rows: one warp per row walking the tensors, compiled for
  // 8 resident blocks per SM (measured 8192 rows: 79 us; (row, tensor) warps 116 us).  (VW 8, 1024 threads) would spill.
  const bool rt = FB_FINALIZE_RT && p.nT <= 8 && grid <= 2 * p.num_sms && !(vw == 8 && p.nT > 4);
  switch (vw) {
    case 1: return launch_finalize_vw<1, PREC>(p, f, g, w, grid, st, rt);
    case 2: return launch_finalize_vw<2, PREC>(p, f, g, w, grid, st, rt);
    case 3: return launch_finalize_vw<3, PREC>(p, f, g, w, grid, st, rt);
    case 4: return launch_finalize_vw<4, PREC>(p, f, g, w, grid, st, rt);
    case 8:
      if constexpr (PREC == FOCAL_PREC_BF16) return launch_finalize_vw<8, PREC>(p, f, g, w, grid, st, rt);
      break;
  }
  return FOCAL_ESHAPE;
}
int launch_finalize_fast(int vw, const Plan& p, const FeatPtrs& f, const GradPtrs& g, const uint8_t* w, int grid,
                         cudaStream_t st) {
  return p.prec == FOCAL_PREC_FP32 ? launch_finalize_fast_p<FOCAL_PREC_FP32>(vw, p, f, g, w, grid, st)
                                   : launch_finalize_fast_p<FOCAL_PREC_BF16>(vw, p, f, g, w, grid, st);
}

// ---- second-generation row kernels (one warp per (row, tensor)): D <= 256 on the vectorised path
bool use_row_v2(int vw) {
  static const int forced_v1 = [] {
    const char* e = std::getenv("FOCAL_B200_ROW_KERNELS");      // "v1": the first-generation kernels (A/B measurements)
    return (e && std::strcmp(e, "v1") == 0) ? 1 : 0;
  }();
  return !forced_v1 && vw >= 1 && vw <= 4;
}
template <int VW, int PREC>
int launch_prologue_v2_vw(const Plan& p, const FeatPtrs& f, const PeerWs& pw, uint8_t* w, int grid, int fuse, cudaStream_t st) {
  const size_t smem = row_v2_smem_bytes(p.nT, p.D);
  if (int rc = ensure_dyn_smem(prologue_v2_kernel<VW, PREC>, smem, "cudaFuncSetAttribute(prologue_v2_kernel)")) return rc;
  prologue_v2_kernel<VW, PREC><<<grid, 128 * p.nT, smem, st>>>(p, f, pw, w, fuse);
  return cuda_ok("prologue_v2_kernel");
}
template <int VW, int PREC>
int launch_finalize_v2_vw(const Plan& p, const FeatPtrs& f, const GradPtrs& g, const uint8_t* w, int grid, cudaStream_t st) {
  const size_t smem = row_v2_smem_bytes(p.nT, p.D);
  if (int rc = ensure_dyn_smem(finalize_v2_kernel<VW, PREC>, smem, "cudaFuncSetAttribute(finalize_v2_kernel)")) return rc;
  finalize_v2_kernel<VW, PREC><<<grid, 128 * p.nT, smem, st>>>(p, f, g, w);
  return cuda_ok("finalize_v2_kernel");
}
#define FB_V2_DISPATCH(FN, ...)                                                                    \
  do {                                                                                             \
    const bool fp32 = p.prec == FOCAL_PREC_FP32;                                                   \
    switch (vw) {                                                                                  \
      case 1: return fp32 ? FN<1, FOCAL_PREC_FP32>(__VA_ARGS__) : FN<1, FOCAL_PREC_BF16>(__VA_ARGS__); \
      case 2: return fp32 ? FN<2, FOCAL_PREC_FP32>(__VA_ARGS__) : FN<2, FOCAL_PREC_BF16>(__VA_ARGS__); \
      case 3: return fp32 ? FN<3, FOCAL_PREC_FP32>(__VA_ARGS__) : FN<3, FOCAL_PREC_BF16>(__VA_ARGS__); \
      case 4: return fp32 ? FN<4, FOCAL_PREC_FP32>(__VA_ARGS__) : FN<4, FOCAL_PREC_BF16>(__VA_ARGS__); \
    }                                                                                              \
    return FOCAL_ESHAPE;                                                                           \
  } while (0)
int launch_prologue_v2(int vw, const Plan& p, const FeatPtrs& f, const PeerWs& pw, uint8_t* w, int grid, int fuse,
                       cudaStream_t st) {
  FB_V2_DISPATCH(launch_prologue_v2_vw, p, f, pw, w, grid, fuse, st);
}
int launch_finalize_v2(int vw, const Plan& p, const FeatPtrs& f, const GradPtrs& g, const uint8_t* w, int grid,
                       cudaStream_t st) {
  FB_V2_DISPATCH(launch_finalize_v2_vw, p, f, g, w, grid, st);
}
#undef FB_V2_DISPATCH

// ---- third-generation row kernels (one warp per (sequence, tensor)); plan.h decides when they apply (Plan::rowgen)
size_t row_v3_smem(const Plan& p) {
  return ((size_t)p.seqb * p.nT * p.S * p.d + (size_t)p.seqb * p.nT * p.S + (size_t)p.seqb * p.nT) * sizeof(float);
}
template <int S, int NQ, int PREC>
int launch_prologue_v3_t(const Plan& p, const FeatPtrs& f, const PeerWs& pw, uint8_t* w, int pad_blocks, cudaStream_t st) {
  const size_t smem = row_v3_smem(p);
  if (int rc = ensure_dyn_smem(prologue_v3_kernel<S, NQ, PREC>, smem, "cudaFuncSetAttribute(prologue_v3_kernel)")) return rc;
  // row shards: the launch can be replicated, the replicas sharing the peers among them (see the kernel).  Measured
  // at 4 GPUs (profiles/r2_prologue_dbg_4.txt): 4 replicas 79 us, 1 replica 60 us -- so the default is 1; the knob stays
  // for experiments.
  int nrep = 1;
  if (pw.world > 1) {
    static const int forced = [] { const char* e = std::getenv("FOCAL_B200_PROLOGUE_REPLICAS"); return e ? std::atoi(e) : 0; }();
    nrep = forced > 0 ? forced : 1;
    while (nrep > 1 && ((long)p.nblk1 * nrep > 4L * p.num_sms || pw.world % nrep)) nrep /= 2;
  }
  launch_pdl(prologue_v3_kernel<S, NQ, PREC>, dim3(p.nblk1 + pad_blocks, nrep), dim3(32 * p.seqb * p.nT), smem, st, p, f, pw, w);
  return cuda_ok("prologue_v3_kernel");
}
template <int S, int NQ, int PREC>
int launch_finalize_v3_t(const Plan& p, const FeatPtrs& f, const GradPtrs& g, const PeerWs& pw, uint8_t* w, float* loss5,
                         cudaStream_t st) {
  const size_t smem = row_v3_smem(p);
  if (int rc = ensure_dyn_smem(finalize_v3_kernel<S, NQ, PREC>, smem, "cudaFuncSetAttribute(finalize_v3_kernel)")) return rc;
  const int grid = (p.seq1 - p.seq0 + p.seqb - 1) / p.seqb + 1;          // + the block that reduces the loss partials
  launch_pdl(finalize_v3_kernel<S, NQ, PREC>, dim3(grid), dim3(32 * p.seqb * p.nT), smem, st, p, f, g, pw, w, loss5,
             lse_blocks(p, 0), temporal_degenerate(p) ? 1 : 0);
  return cuda_ok("finalize_v3_kernel");
}
#define FB_V3_CASE(FN, SS, QQ, ...)                                                                   \
  if (p.S == SS && p.nq == QQ)                                                                        \
    return p.prec == FOCAL_PREC_FP32 ? FN<SS, QQ, FOCAL_PREC_FP32>(__VA_ARGS__) : FN<SS, QQ, FOCAL_PREC_BF16>(__VA_ARGS__)
#ifdef FB_FAST_BUILD                  // experiment builds (tools/variant_bench.py) only instantiate the headline shapes
#define FB_V3_DISPATCH(FN, ...)                                                                       \
  do {                                                                                                \
    FB_V3_CASE(FN, 4, 4, __VA_ARGS__);                                                                \
    return FOCAL_ESHAPE;                                                                              \
  } while (0)
#else
#define FB_V3_DISPATCH(FN, ...)                                                                       \
  do {                                                                                                \
    FB_V3_CASE(FN, 4, 4, __VA_ARGS__); FB_V3_CASE(FN, 4, 2, __VA_ARGS__);                             \
    FB_V3_CASE(FN, 4, 1, __VA_ARGS__); FB_V3_CASE(FN, 4, 3, __VA_ARGS__);                             \
    FB_V3_CASE(FN, 2, 1, __VA_ARGS__); FB_V3_CASE(FN, 2, 2, __VA_ARGS__);                             \
    FB_V3_CASE(FN, 1, 1, __VA_ARGS__);                                                                \
    return FOCAL_ESHAPE;                                                                              \
  } while (0)
#endif
int launch_prologue_v3(const Plan& p, const FeatPtrs& f, const PeerWs& pw, uint8_t* w, int pad_blocks, cudaStream_t st) {
  FB_V3_DISPATCH(launch_prologue_v3_t, p, f, pw, w, pad_blocks, st);
}
int launch_finalize_v3(const Plan& p, const FeatPtrs& f, const GradPtrs& g, const PeerWs& pw, uint8_t* w, float* loss5,
                       cudaStream_t st) {
  FB_V3_DISPATCH(launch_finalize_v3_t, p, f, g, pw, w, loss5, st);
}
#undef FB_V3_DISPATCH
#undef FB_V3_CASE

// local_rows: the caller's tensors start at the first owned row; the kernels index rows globally, so hand them the
// (virtual) address of row 0 -- only owned rows are ever dereferenced.
int fill_feats(const Plan& p, const float* const* feats, FeatPtrs& f) {
  if (p.indirect) {                       // the kernels read the table focal_b200_set_ptrs filled
    for (int t = 0; t < kMaxT; ++t) f.x[t] = nullptr;
    return FOCAL_OK;
  }
  if (!feats) return FOCAL_EINVAL;
  const ptrdiff_t shift = p.local_rows ? (ptrdiff_t)p.seq0 * p.S * p.D : 0;
  for (int t = 0; t < p.nT; ++t) {
    if (!feats[t] || (reinterpret_cast<uintptr_t>(feats[t]) & 15)) return FOCAL_EINVAL;
    f.x[t] = feats[t] - shift;
  }
  for (int t = p.nT; t < kMaxT; ++t) f.x[t] = nullptr;
  return FOCAL_OK;
}
int fill_grads(const Plan& p, float* const* grads, GradPtrs& g) {
  if (p.indirect) {
    for (int t = 0; t < kMaxT; ++t) g.g[t] = nullptr;
    return FOCAL_OK;
  }
  if (!grads) return FOCAL_EINVAL;
  const ptrdiff_t shift = p.local_rows ? (ptrdiff_t)p.seq0 * p.S * p.D : 0;
  for (int t = 0; t < kMaxT; ++t) g.g[t] = nullptr;
  for (int t = 0; t < p.nT; ++t) {
    if (!grads[t] || (reinterpret_cast<uintptr_t>(grads[t]) & 15)) return FOCAL_EINVAL;
    g.g[t] = grads[t] - shift;
  }
  return FOCAL_OK;
}
PeerWs solo(void* ws) {
  PeerWs pw{};
  pw.rank = 0; pw.world = 1;
  pw.ws[0] = static_cast<uint8_t*>(ws);
  return pw;
}

int do_prologue(const Plan& p, int no_private, const FeatPtrs& f, const PeerWs& pw, uint8_t* w, cudaStream_t st,
                bool zero_pads = true) {
  int rc;
  const bool pads = zero_pads && (p.bpad != p.b || p.Bpad != p.Bt || p.Sp != p.S);
  const bool pads_in_prologue = pads && p.rowgen == 3 && pw.world == 1;     // extra blocks of prologue_v3 do it
  if (pads && !pads_in_prologue) {
    zero_pad_kernel<<<64, 256, 0, st>>>(p, w);
    if ((rc = cuda_ok("zero_pad_kernel"))) return rc;
  }
  const int vw = fast_row_vw(p, no_private);
  bool fused_intra = false;
  if (p.rowgen == 3) {
    fused_intra = true;                 // S == 1: the temporal term is degenerate, nothing to fuse
    if ((rc = launch_prologue_v3(p, f, pw, w, pads_in_prologue ? 8 : 0, st))) return rc;
  } else if (vw) {
    fused_intra = (p.S == 2 || p.S == 4);
    const size_t smem = ((size_t)4 * p.nT * p.D + 4 * 2 * kMaxT + 4 * kMaxT) * sizeof(float);
    // finalize_v2 reads the squared norms prologue_v2 stores, so the two generations are only used as a pair
    if (use_row_v2(vw) && p.nT <= 8) rc = launch_prologue_v2(vw, p, f, pw, w, p.nblk1, fused_intra ? 1 : 0, st);
    else rc = launch_prologue_fast(vw, p, f, pw, w, smem, p.nblk1, fused_intra ? 1 : 0, st);
    if (rc) return rc;
  } else {
    if (pw.world > 1 || p.local_rows) return FOCAL_ESHAPE;      // the generic row kernels have no peer path
    const size_t smem = (size_t)kRowsPerBlock * p.nT * p.D * sizeof(float);
    if ((rc = ensure_dyn_smem(prologue_kernel, smem, "cudaFuncSetAttribute(prologue_kernel)"))) return rc;
    prologue_kernel<<<p.nblk1, 32 * kRowsPerBlock, smem, st>>>(p, f, w);
    if ((rc = cuda_ok("prologue_kernel"))) return rc;
  }
  if (!fused_intra && (p.terms & FOCAL_TERM_TEMPORAL) && !temporal_degenerate(p)) {
    if (pw.world > 1 || p.local_rows) return FOCAL_ESHAPE;
    const long warps = (long)p.nT * p.b;
    intra_kernel<<<(unsigned)((warps + 3) / 4), 128, 0, st>>>(p, f, w);
    if ((rc = cuda_ok("intra_kernel"))) return rc;
  }
  return FOCAL_OK;
}

int do_finalize(const Plan& p, int no_private, const float* const* feats, float* const* grads, const PeerWs& pw,
                uint8_t* w, float* loss5, cudaStream_t st) {
  int rc;
  if (p.need_grad) {
    FeatPtrs f;
    if ((rc = fill_feats(p, feats, f))) return rc;
    GradPtrs g;
    if ((rc = fill_grads(p, grads, g))) return rc;
    const int rows = (p.seq1 - p.seq0) * p.S;
    const int vw = (p.S == 1 || p.S == 2 || p.S == 4) ? fast_row_vw(p, no_private) : 0;
    if (p.rowgen == 3) {
      return launch_finalize_v3(p, f, g, pw, w, loss5, st);           // its last block is the loss reduction
    } else if (vw) {
      if (use_row_v2(vw) && p.nT <= 8) rc = launch_finalize_v2(vw, p, f, g, w, (rows + 3) / 4, st);
      else rc = launch_finalize_fast(vw, p, f, g, w, (rows + 3) / 4, st);
      if (rc) return rc;
    } else {
      const size_t smem = (size_t)kRowsPerBlock * (2 * p.nT + 1) * p.D * sizeof(float);
      if ((rc = ensure_dyn_smem(finalize_kernel, smem, "cudaFuncSetAttribute(finalize_kernel)"))) return rc;
      finalize_kernel<<<(rows + kRowsPerBlock - 1) / kRowsPerBlock, 32 * kRowsPerBlock, smem, st>>>(p, f, g, w);
      if ((rc = cuda_ok("finalize_kernel"))) return rc;
    }
  }
  loss_reduce_kernel<<<1, 256, 0, st>>>(p, pw, w, loss5, lse_blocks(p, 0), temporal_degenerate(p) ? 1 : 0);
  return cuda_ok("loss_reduce_kernel");
}

}  // namespace

extern "C" {

int focal_b200_abi_version(void) { return FOCAL_B200_ABI_VERSION; }

const char* focal_b200_strerror(int code) {
  switch (code) {
    case FOCAL_OK: return "ok";
    case FOCAL_EINVAL: return "invalid argument";
    case FOCAL_ESHAPE:
      return "unsupported shape: need B % S == 0, S <= 32, 2 <= D <= 512 (fp32 mode: <= 256), 1 <= M <= 8, "
             "temperature >= 0.016";
    case FOCAL_ECUDA: return "CUDA error (see stderr)";
    case FOCAL_EWORKSPACE: return "workspace too small or not 1024-byte aligned";
  }
  return "unknown error";
}

int focal_b200_workspace_info(const FocalCfg* cfg, FocalWsInfo* info) {
  Plan p;
  int rc = make_plan(cfg, p);
  if (rc) return rc;
  if (!info) return FOCAL_EINVAL;
  info->total_bytes = p.total_bytes;
  info->b = p.b; info->bpad = p.bpad; info->Bpad = p.Bpad; info->n_problems = p.nProb; info->n_ops = p.nOps;
  info->kb_full = p.kbFull;
  info->rowsum_off = p.rsum_off;
  info->rowsum_bytes = (size_t)p.nProb * p.S * 2 * p.bpad * 4;
  info->cnt_off = p.cnt_off;
  info->cnt_bytes = (size_t)p.nT * p.bpad * 4;
  info->mintra_off = p.mintra_off;
  info->lossparts_off = p.lossd_off;
  info->dz_off = p.dz_off; info->dz_bytes = p.dz_bytes;
  info->dx_off = p.dx_off; info->dx_bytes = p.dx_bytes;
  info->cnt_piece_stride = p.cnt2_delta;
  info->n_pieces_nce = p.np_nce; info->n_pieces_tmp = p.np_tmp;
  return FOCAL_OK;
}

int focal_b200_prologue(const FocalCfg* cfg, const float* const* feats, void* ws, size_t ws_bytes, void* stream) {
  Plan p;
  int rc = make_plan(cfg, p);
  if (rc) return rc;
  if (p.local_rows) return FOCAL_EINVAL;               // local_rows belongs to focal_b200_loss_sharded
  if ((rc = check_ws(p, ws, ws_bytes))) return rc;
  FeatPtrs f;
  if ((rc = fill_feats(p, feats, f))) return rc;
  return do_prologue(p, cfg->no_private, f, solo(ws), static_cast<uint8_t*>(ws), static_cast<cudaStream_t>(stream));
}

int focal_b200_nce_rowsum(const FocalCfg* cfg, void* ws, size_t ws_bytes, void* stream) {
  Plan p;
  int rc = make_plan(cfg, p);
  if (rc) return rc;
  if ((rc = check_ws(p, ws, ws_bytes))) return rc;
  if (!(p.terms & FOCAL_TERM_NCE)) return FOCAL_OK;
  return launch_nce<NCE_FWD>(p, static_cast<uint8_t*>(ws), static_cast<cudaStream_t>(stream));
}

int focal_b200_nce_lse(const FocalCfg* cfg, void* ws, size_t ws_bytes, int all_rows, void* stream) {
  Plan p;
  int rc = make_plan(cfg, p);
  if (rc) return rc;
  if ((rc = check_ws(p, ws, ws_bytes))) return rc;
  if (!(p.terms & FOCAL_TERM_NCE)) return FOCAL_OK;
  launch_pdl(nce_lse_kernel, dim3(lse_blocks(p, all_rows)), dim3(256), 0, static_cast<cudaStream_t>(stream), p, solo(ws),
             static_cast<uint8_t*>(ws), all_rows);
  return cuda_ok("nce_lse_kernel");
}

int focal_b200_nce_grad(const FocalCfg* cfg, void* ws, size_t ws_bytes, void* stream) {
  Plan p;
  int rc = make_plan(cfg, p);
  if (rc) return rc;
  if ((rc = check_ws(p, ws, ws_bytes))) return rc;
  if (!(p.terms & FOCAL_TERM_NCE) || !p.need_grad) return FOCAL_OK;
  return launch_nce<NCE_BWD>(p, static_cast<uint8_t*>(ws), static_cast<cudaStream_t>(stream));
}

int focal_b200_temporal(const FocalCfg* cfg, void* ws, size_t ws_bytes, void* stream) {
  Plan p;
  int rc = make_plan(cfg, p);
  if (rc) return rc;
  if ((rc = check_ws(p, ws, ws_bytes))) return rc;
  if (!(p.terms & FOCAL_TERM_TEMPORAL) || temporal_degenerate(p)) return FOCAL_OK;
  uint8_t* w = static_cast<uint8_t*>(ws);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  return p.need_grad ? launch_temporal<TMP_BWD>(p, w, st, p.grid_tmp) : launch_temporal<TMP_FWD>(p, w, st, p.grid_tmp);
}

int focal_b200_finalize(const FocalCfg* cfg, const float* const* feats, void* ws, size_t ws_bytes, float* loss5,
                        float* const* grads, void* stream) {
  Plan p;
  int rc = make_plan(cfg, p);
  if (rc) return rc;
  if (p.local_rows) return FOCAL_EINVAL;
  if ((rc = check_ws(p, ws, ws_bytes))) return rc;
  if (!loss5 && !p.indirect) return FOCAL_EINVAL;
  return do_finalize(p, cfg->no_private, feats, grads, solo(ws), static_cast<uint8_t*>(ws), loss5,
                     static_cast<cudaStream_t>(stream));
}

int focal_b200_set_ptrs(const FocalCfg* cfg, void* ws, size_t ws_bytes, const float* const* feats, float* loss5,
                        float* const* grads, void* stream) {
  Plan p;
  int rc = make_plan(cfg, p);
  if (rc) return rc;
  if ((rc = check_ws(p, ws, ws_bytes))) return rc;
  if (!p.indirect || !feats || !loss5 || (p.need_grad && !grads)) return FOCAL_EINVAL;
  const ptrdiff_t shift = p.local_rows ? (ptrdiff_t)p.seq0 * p.S * p.D : 0;
  PtrTable t{};
  for (int i = 0; i < p.nT; ++i) {
    if (!feats[i] || (reinterpret_cast<uintptr_t>(feats[i]) & 15)) return FOCAL_EINVAL;
    t.x[i] = feats[i] - shift;
    if (p.need_grad) {
      if (!grads[i] || (reinterpret_cast<uintptr_t>(grads[i]) & 15)) return FOCAL_EINVAL;
      t.g[i] = grads[i] - shift;
    }
  }
  t.loss5 = loss5;
  set_ptrs_kernel<<<1, 32, 0, static_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<PtrTable*>(static_cast<uint8_t*>(ws) + p.ptrs_off), t);
  return cuda_ok("set_ptrs_kernel");
}

int focal_b200_loss(const FocalCfg* cfg, const float* const* feats, void* ws, size_t ws_bytes, float* loss5,
                    float* const* grads, void* stream) {
  int rc;
  if ((rc = focal_b200_prologue(cfg, feats, ws, ws_bytes, stream))) return rc;
  if ((rc = focal_b200_nce_rowsum(cfg, ws, ws_bytes, stream))) return rc;
  if ((rc = focal_b200_nce_lse(cfg, ws, ws_bytes, 0, stream))) return rc;
  if ((rc = focal_b200_nce_grad(cfg, ws, ws_bytes, stream))) return rc;
  if ((rc = focal_b200_temporal(cfg, ws, ws_bytes, stream))) return rc;
  return focal_b200_finalize(cfg, feats, ws, ws_bytes, loss5, grads, stream);
}

// ---------------------------------------------------------------------------------------------------------
// row-sharded path over NVLink peer memory
// ---------------------------------------------------------------------------------------------------------
int focal_b200_peer_alloc(size_t bytes, void** ptr, unsigned char handle[64]) {
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "handle size");
  if (!ptr || !handle || !bytes) return FOCAL_EINVAL;
  void* d = nullptr;
  if (cudaMalloc(&d, bytes) != cudaSuccess) return cuda_ok("cudaMalloc(peer workspace)");
  if (cudaMemset(d, 0, bytes) != cudaSuccess || cudaDeviceSynchronize() != cudaSuccess) {
    cudaFree(d);
    return cuda_ok("cudaMemset(peer workspace)");
  }
  cudaIpcMemHandle_t h;
  if (cudaIpcGetMemHandle(&h, d) != cudaSuccess) {
    cudaFree(d);
    return cuda_ok("cudaIpcGetMemHandle");
  }
  std::memcpy(handle, &h, 64);
  *ptr = d;
  return FOCAL_OK;
}

int focal_b200_peer_open(const unsigned char handle[64], void** ptr) {
  if (!ptr || !handle) return FOCAL_EINVAL;
  cudaIpcMemHandle_t h;
  std::memcpy(&h, handle, 64);
  void* d = nullptr;
  if (cudaIpcOpenMemHandle(&d, h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) return cuda_ok("cudaIpcOpenMemHandle");
  *ptr = d;
  return FOCAL_OK;
}

int focal_b200_peer_close(void* ptr) {
  if (!ptr) return FOCAL_EINVAL;
  return cudaIpcCloseMemHandle(ptr) == cudaSuccess ? FOCAL_OK : cuda_ok("cudaIpcCloseMemHandle");
}

int focal_b200_peer_free(void* ptr) {
  if (!ptr) return FOCAL_EINVAL;
  return cudaFree(ptr) == cudaSuccess ? FOCAL_OK : cuda_ok("cudaFree(peer workspace)");
}

namespace {
// Diagnostics (FOCAL_B200_STAGE_TIMES=1 in the environment, eager launches only -- tools/shard_stage_times.py): CUDA
// events between the launches of focal_b200_loss_sharded, read back with focal_b200_debug_stage_times.
struct StageTimer {
  bool on = false, init = false;
  cudaEvent_t ev[8];
  int n = 0;
  void begin() {
    if (!init) {
      const char* e = std::getenv("FOCAL_B200_STAGE_TIMES");
      on = e && e[0] == '1';
      if (on)
        for (auto& x : ev) cudaEventCreate(&x);
      init = true;
    }
    n = 0;
  }
  void mark(cudaStream_t st) {
    if (!on || n >= 8) return;
    cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(st, &cs) != cudaSuccess || cs != cudaStreamCaptureStatusNone) return;
    cudaEventRecord(ev[n++], st);
  }
};
StageTimer g_stage_timer;

// Do two ranks of this peer table keep their workspace on the same device (several ranks emulated on one GPU)?
bool ranks_share_a_device(const FocalPeers* peers) {
  static FocalPeers seen{};
  static bool seen_valid = false, seen_result = false;
  if (seen_valid && std::memcmp(&seen, peers, sizeof(FocalPeers)) == 0) return seen_result;
  bool shared = false;
  int dev[FOCAL_MAX_PEERS];
  for (int r = 0; r < peers->world; ++r) {
    cudaPointerAttributes a{};
    dev[r] = (cudaPointerGetAttributes(&a, peers->ws[r]) == cudaSuccess) ? a.device : -1 - r;
    for (int q = 0; q < r; ++q) shared = shared || dev[q] == dev[r];
  }
  cudaGetLastError();
  seen = *peers; seen_valid = true; seen_result = shared;
  return shared;
}
}  // namespace

int focal_b200_loss_sharded(const FocalCfg* cfg, const float* const* feats, const FocalPeers* peers, size_t ws_bytes,
                            float* loss5, float* const* grads, void* stream) {
  Plan p;
  int rc = make_plan(cfg, p);
  if (rc) return rc;
  if (!peers || (!loss5 && !p.indirect) || !p.local_rows) return FOCAL_EINVAL;
  if (peers->world < 1 || peers->world > kMaxPeers || peers->rank < 0 || peers->rank >= peers->world) return FOCAL_EINVAL;
  PeerWs pw{};
  pw.rank = peers->rank; pw.world = peers->world;
  // the multicast mapping only serves the kernels that know it (prologue_v3, nce_lse); the older row kernels store per peer
  pw.mc = (peers->world > 1 && p.rowgen == 3) ? static_cast<uint8_t*>(peers->mc) : nullptr;
  for (int r = 0; r < pw.world; ++r) {
    if ((rc = check_ws(p, peers->ws[r], ws_bytes))) return rc;
    pw.ws[r] = static_cast<uint8_t*>(peers->ws[r]);
  }
  uint8_t* w = pw.ws[pw.rank];
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  FeatPtrs f;
  if ((rc = fill_feats(p, feats, f))) return rc;
  StageTimer& tm = g_stage_timer;
  tm.begin();
  tm.mark(st);
  // phase 1: operands of the owned rows -> every workspace.  The prologue only counts the epoch; block 0 of the NEXT launch
  // (stream order: every store of the prologue is complete) announces it to the peers, and every block of that launch
  // waits for all announcements before it reads operands.  (No zero_pad launch: workspaces from focal_b200_peer_alloc
  // start zeroed and nobody ever writes a padding row.)
  if ((rc = do_prologue(p, cfg->no_private, f, pw, w, st, /*zero_pads=*/false))) return rc;
  tm.mark(st);
  // several ranks on one device: announce / wait become one-block launches of their own (see peer_wait_kernel)
  const bool multi = pw.world > 1;
  const bool split = multi && ranks_share_a_device(peers);
  auto sync_launch = [&](bool announce, bool wait) -> int {
    if (!split) return FOCAL_OK;
    peer_wait_kernel<<<1, 32, 0, st>>>(p, pw, announce ? 1 : 0, wait ? 1 : 0);
    return cuda_ok("peer_wait_kernel");
  };
  auto sync_sel = [&](bool announce, bool wait) {
    ProbSel s{};
    if (multi && !split) {
      s.peer_wait = wait ? pw.world : 0;
      s.ann_world = announce ? pw.world : 0;
      s.ann_rank = pw.rank;
      for (int r = 0; r < pw.world; ++r) s.peer_ws[r] = pw.ws[r];
    }
    return s;
  };
  const bool nce = (p.terms & FOCAL_TERM_NCE) != 0;
  const bool tmp = (p.terms & FOCAL_TERM_TEMPORAL) && !temporal_degenerate(p);
  // phase 2: row sums of the owned rows -> every workspace (counted by nce_lse, announced by the launch after it).  The
  // temporal launch needs nothing from the peers beyond phase 1, so it runs between those stores and the launch that
  // waits for them.
  if (nce || tmp) {
    if ((rc = sync_launch(true, true))) return rc;
  }
  bool pending = false;               // a counted epoch that has not been announced yet
  if (nce) {
    if ((rc = launch_nce<NCE_FWD>(p, w, st, sync_sel(true, true)))) return rc;
    tm.mark(st);
    launch_pdl(nce_lse_kernel, dim3(lse_blocks(p, 0)), dim3(256), 0, st, p, pw, w, 0);
    if ((rc = cuda_ok("nce_lse_kernel"))) return rc;
    pending = multi;
    tm.mark(st);
  }
  if (tmp) {
    if (pending && (rc = sync_launch(true, false))) return rc;
    const ProbSel s = nce ? sync_sel(pending, false) : sync_sel(true, true);
    rc = p.need_grad ? launch_temporal<TMP_BWD>(p, w, st, p.grid_tmp, s) : launch_temporal<TMP_FWD>(p, w, st, p.grid_tmp, s);
    if (rc) return rc;
    pending = false;
    tm.mark(st);
  }
  if (nce && p.need_grad) {
    if ((rc = sync_launch(pending, true))) return rc;
    if ((rc = launch_nce<NCE_BWD>(p, w, st, sync_sel(pending, true)))) return rc;
    pending = false;
    tm.mark(st);
  }
  // phase 3: gradients of the owned rows; loss partials all-reduced by the last block of the finalize launch (third barrier)
  rc = do_finalize(p, cfg->no_private, feats, grads, pw, w, loss5, st);
  tm.mark(st);
  return rc;
}

// Diagnostics: milliseconds between the stage marks of the last eager focal_b200_loss_sharded call of this process
// (prologue, nce_rowsum, nce_lse, temporal, nce_grad, finalize for the full loss); synchronises.  Returns the number of
// intervals written, 0 when FOCAL_B200_STAGE_TIMES is not set.
int focal_b200_debug_stage_times(float* out_ms, int n) {
  StageTimer& tm = g_stage_timer;
  if (!tm.on || !out_ms || tm.n < 2) return 0;
  if (cudaEventSynchronize(tm.ev[tm.n - 1]) != cudaSuccess) return 0;
  int k = 0;
  for (; k + 1 < tm.n && k < n; ++k)
    if (cudaEventElapsedTime(&out_ms[k], tm.ev[k], tm.ev[k + 1]) != cudaSuccess) return 0;
  return k;
}

}  // extern "C"

// ---------------------------------------------------------------------------------------------------------
// frequency-domain input stage
// ---------------------------------------------------------------------------------------------------------
extern "C" int focal_b200_spectrum_rotate(const float* in, float* out, long long n_bc, int plane, int interleaved,
                                          float cos_angle, float sin_angle, void* stream) {
  if (!in || !out || n_bc <= 0 || plane <= 0 || (plane & 3)) return FOCAL_EINVAL;
  if ((reinterpret_cast<uintptr_t>(in) & 15) || (reinterpret_cast<uintptr_t>(out) & 15)) return FOCAL_EINVAL;
  const long long vecs = n_bc * (plane >> 2);
  long long blocks = (vecs + 255) / 256;
  const long long cap = (long long)device_sms() * 16;          // grid-stride: a few waves of 256-thread blocks
  if (blocks > cap) blocks = cap;
  fb::spectrum_rotate_kernel<<<(unsigned)blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      in, out, n_bc, plane, interleaved, cos_angle, sin_angle);
  return cuda_ok("spectrum_rotate_kernel");
}

// ---------------------------------------------------------------------------------------------------------
// k-nearest-neighbour evaluation
// ---------------------------------------------------------------------------------------------------------
extern "C" int focal_b200_knn_predict(const float* train, const int32_t* labels, int n_train, const float* query,
                                      int n_query, int dim, int k, int n_classes, float* dist_ws, int32_t* out,
                                      int32_t* neighbours, void* stream) {
  if (!train || !labels || !query || !dist_ws || !out) return FOCAL_EINVAL;
  if (n_train <= 0 || n_query <= 0 || dim <= 0 || k <= 0) return FOCAL_EINVAL;
  if (k > fb::kKnnMaxK || k > n_train || n_classes < 1 || n_classes > 32) return FOCAL_ESHAPE;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  dim3 grid((n_train + 63) / 64, (n_query + 63) / 64);
  fb::knn_dist_kernel<<<grid, 256, 0, st>>>(query, train, n_query, n_train, dim, dist_ws);
  int rc = cuda_ok("knn_dist_kernel");
  if (rc) return rc;
  fb::knn_select_kernel<<<(n_query + 3) / 4, 128, 0, st>>>(dist_ws, labels, n_query, n_train, k, n_classes, out, neighbours);
  return cuda_ok("knn_select_kernel");
}

#ifdef FB_TRACE
// experiment builds only: copy the pipeline trace of the last Gram launch to the host (tools/trace_gram.py)
extern "C" int focal_b200_debug_trace(long long* host_dst, size_t n) {
  const size_t have = sizeof(fb::fb_trace_buf) / sizeof(long long);
  if (!host_dst || n > have) return FOCAL_EINVAL;
  if (cudaDeviceSynchronize() != cudaSuccess) return FOCAL_ECUDA;
  return cudaMemcpyFromSymbol(host_dst, fb::fb_trace_buf, n * sizeof(long long)) == cudaSuccess ? FOCAL_OK : FOCAL_ECUDA;
}
extern "C" int focal_b200_debug_trace_clear(void) {
  void* ptr = nullptr;
  if (cudaGetSymbolAddress(&ptr, fb::fb_trace_buf) != cudaSuccess) return FOCAL_ECUDA;
  return cudaMemset(ptr, 0, sizeof(fb::fb_trace_buf)) == cudaSuccess ? FOCAL_OK : FOCAL_ECUDA;
}
#endif
