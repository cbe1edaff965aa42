// C ABI of the B200-native FOCAL loss hot path (see include/focal_b200.h).
#include <cstdio>
#include <cstring>
#include <cuda_runtime.h>

#include "../../include/focal_b200.h"
#include "gram_kernel.cuh"
#include "plan.h"
#include "row_kernels.cuh"
#include "row_kernels_fast.cuh"

using namespace fb;

namespace {

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

int make_plan(const FocalCfg* cfg, Plan& p) {
  if (!cfg) return FOCAL_EINVAL;
  const int sms = cfg->num_sms > 0 ? cfg->num_sms : device_sms();
  return build_plan(*cfg, p, sms);
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

template <int MODE, int KB, int SEQ>
int launch_gram(const Plan& p, const ProbSel& sel, uint8_t* ws, cudaStream_t st, int grid) {
  using G = GramCfg<MODE, KB, SEQ>;
  static_assert(G::BN == tile_bn(KB), "plan.h and gram_kernel.cuh must agree on the column tile");
  using L = GramSmem<G::BN, KB, G::NB>;
  auto kfn = gram_kernel<MODE, KB, SEQ>;
  static bool configured = false;     // per instantiation; the attribute is sticky per context
  if (!configured) {
    if (cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L::kDynamic) != cudaSuccess)
      return cuda_ok("cudaFuncSetAttribute(gram_kernel)");
    configured = true;
  }
  if (grid <= 0) return FOCAL_OK;                                 // persistent: at most one CTA per SM (plan.h)
  kfn<<<grid, G::kThreads, L::kDynamic, st>>>(p, sel, ws);
  return cuda_ok("gram_kernel launch");
}

template <int MODE, int SEQ>
int launch_gram_kb(const Plan& p, const ProbSel& sel, uint8_t* ws, cudaStream_t st, int kb, int grid) {
  switch (kb) {
#ifndef FB_FAST_BUILD
    case 1: return launch_gram<MODE, 1, SEQ>(p, sel, ws, st, grid);
    case 3: return launch_gram<MODE, 3, SEQ>(p, sel, ws, st, grid);
#endif
    case 2: return launch_gram<MODE, 2, SEQ>(p, sel, ws, st, grid);
    case 4: return launch_gram<MODE, 4, SEQ>(p, sel, ws, st, grid);
  }
  return FOCAL_ESHAPE;
}

// temporal launches: operand width 1..4 K blocks, or 8 ("wide": 256 < D <= 512)
template <int MODE, int SEQ>
int launch_temporal_kb(const Plan& p, const ProbSel& sel, uint8_t* ws, cudaStream_t st, int grid) {
#ifndef FB_FAST_BUILD
  if (p.kbFull == 8) return launch_gram<MODE, 8, SEQ>(p, sel, ws, st, grid);
#endif
  return launch_gram_kb<MODE, SEQ>(p, sel, ws, st, p.kbFull, grid);
}

template <int MODE>
int launch_temporal(const Plan& p, uint8_t* ws, cudaStream_t st, int grid, int peer_wait = 0) {
  ProbSel sel{};
  sel.peer_wait = peer_wait;
  switch (p.S) {
    case 4: return launch_temporal_kb<MODE, 4>(p, sel, ws, st, grid);
#ifndef FB_FAST_BUILD                 // experiment builds (tools/variant_bench.py) only instantiate the headline shapes
    case 2: return launch_temporal_kb<MODE, 2>(p, sel, ws, st, grid);
    case 8: return launch_temporal_kb<MODE, 8>(p, sel, ws, st, grid);
    case 16: return launch_temporal_kb<MODE, 16>(p, sel, ws, st, grid);
    case 32: return launch_temporal_kb<MODE, 32>(p, sel, ws, st, grid);
#endif
  }
  return FOCAL_ESHAPE;
}

// InfoNCE problems are launched in groups of equal operand width (they differ only under noPrivate, where the
// shared problems use the full D columns and the private ones D/2).
template <int MODE>
int launch_nce(const Plan& p, uint8_t* ws, cudaStream_t st, int peer_wait = 0) {
  for (int kb = 1; kb <= 4; ++kb) {
    ProbSel sel{};
    sel.peer_wait = peer_wait;
    for (int q = 0; q < p.nProb; ++q)
      if (p.ops[p.probs[q].opA].kb == kb) sel.idx[sel.n++] = q;
    if (!sel.n) continue;
    const int rc = launch_gram_kb<MODE, 0>(p, sel, ws, st, kb, p.grid_nce[kb]);
    if (rc) return rc;
  }
  return FOCAL_OK;
}

// lanes own VW = d/32 consecutive columns per half on the vectorised row-kernel path (0 = use the generic kernels)
int fast_row_vw(const Plan& p, int no_private) {
  if (no_private || (p.D & 1) || p.d % 32 || p.d < 32) return 0;
  const int vw = p.d / 32;
  return (vw <= 4 || vw == 8) ? vw : 0;           // D = 64, 128, 192, 256, 512
}

template <int VW>
int launch_prologue_fast_vw(const Plan& p, const FeatPtrs& f, const PeerWs& pw, uint8_t* w, size_t smem, int grid, int fuse,
                            cudaStream_t st) {
  static size_t configured = 0;
  if (smem > 48 * 1024 && smem > configured) {
    if (cudaFuncSetAttribute(prologue_fast_kernel<VW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
      return cuda_ok("cudaFuncSetAttribute(prologue_fast_kernel)");
    configured = smem;
  }
  prologue_fast_kernel<VW><<<grid, 128, smem, st>>>(p, f, pw, w, fuse);
  return cuda_ok("prologue_fast_kernel");
}
int launch_prologue_fast(int vw, const Plan& p, const FeatPtrs& f, const PeerWs& pw, uint8_t* w, size_t smem, int grid,
                         int fuse, cudaStream_t st) {
  switch (vw) {
    case 1: return launch_prologue_fast_vw<1>(p, f, pw, w, smem, grid, fuse, st);
    case 2: return launch_prologue_fast_vw<2>(p, f, pw, w, smem, grid, fuse, st);
    case 3: return launch_prologue_fast_vw<3>(p, f, pw, w, smem, grid, fuse, st);
    case 4: return launch_prologue_fast_vw<4>(p, f, pw, w, smem, grid, fuse, st);
    case 8: return launch_prologue_fast_vw<8>(p, f, pw, w, smem, grid, fuse, st);
  }
  return FOCAL_ESHAPE;
}
#ifndef FB_FINALIZE_RT
#define FB_FINALIZE_RT 1          // 1: one warp per (row, tensor) when the launch is latency-bound; 0: always one warp per row
#endif
template <int VW, int MAXT>
int launch_finalize_rt_t(const Plan& p, const FeatPtrs& f, const GradPtrs& g, const uint8_t* w, int grid, cudaStream_t st) {
  const size_t smem = ((size_t)4 * p.nT * p.D + 4 * 2 * kMaxT) * sizeof(float);
  static size_t configured = 0;
  if (smem > 48 * 1024 && smem > configured) {
    if (cudaFuncSetAttribute(finalize_rt_kernel<VW, MAXT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
      return cuda_ok("cudaFuncSetAttribute(finalize_rt_kernel)");
    configured = smem;
  }
  finalize_rt_kernel<VW, MAXT><<<grid, MAXT > 0 ? 128 * p.nT : 128, smem, st>>>(p, f, g, w);
  return cuda_ok("finalize_rt_kernel");
}
template <int VW>
int launch_finalize_rt_vw(const Plan& p, const FeatPtrs& f, const GradPtrs& g, const uint8_t* w, int grid, cudaStream_t st) {
  // <= 4 tensors (M <= 2): 512-thread blocks, 128 registers per thread available; else 1024-thread blocks
  return p.nT <= 4 ? launch_finalize_rt_t<VW, 4>(p, f, g, w, grid, st) : launch_finalize_rt_t<VW, 8>(p, f, g, w, grid, st);
}
int launch_finalize_fast(int vw, const Plan& p, const FeatPtrs& f, const GradPtrs& g, const uint8_t* w, int grid,
                         cudaStream_t st) {
  // Few rows (a row shard): the launch is one wave of blocks and its time is the dependent chain of one warp -> one warp
  // per (row, tensor) (measured 1024 rows: 41 -> 32 us).  Many rows: one warp per row walking the tensors, compiled for
  // 8 resident blocks per SM (measured 8192 rows: 79 us; (row, tensor) warps 116 us).  (VW 8, 1024 threads) would spill.
  const bool rt = FB_FINALIZE_RT && p.nT <= 8 && grid <= 2 * p.num_sms && !(vw == 8 && p.nT > 4);
  switch (vw) {
    case 1: return rt ? launch_finalize_rt_vw<1>(p, f, g, w, grid, st) : launch_finalize_rt_t<1, 0>(p, f, g, w, grid, st);
    case 2: return rt ? launch_finalize_rt_vw<2>(p, f, g, w, grid, st) : launch_finalize_rt_t<2, 0>(p, f, g, w, grid, st);
    case 3: return rt ? launch_finalize_rt_vw<3>(p, f, g, w, grid, st) : launch_finalize_rt_t<3, 0>(p, f, g, w, grid, st);
    case 4: return rt ? launch_finalize_rt_vw<4>(p, f, g, w, grid, st) : launch_finalize_rt_t<4, 0>(p, f, g, w, grid, st);
    case 8: return rt ? launch_finalize_rt_vw<8>(p, f, g, w, grid, st) : launch_finalize_rt_t<8, 0>(p, f, g, w, grid, st);
  }
  return FOCAL_ESHAPE;
}

// local_rows: the caller's tensors start at the first owned row; the kernels index rows globally, so hand them the
// (virtual) address of row 0 -- only owned rows are ever dereferenced.
int fill_feats(const Plan& p, const float* const* feats, FeatPtrs& f) {
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
  if (zero_pads && (p.bpad != p.b || p.Bpad != p.B)) {
    zero_pad_kernel<<<64, 256, 0, st>>>(p, w);
    if ((rc = cuda_ok("zero_pad_kernel"))) return rc;
  }
  const int vw = fast_row_vw(p, no_private);
  bool fused_intra = false;
  if (vw) {
    fused_intra = (p.S == 2 || p.S == 4);
    const size_t smem = ((size_t)4 * p.nT * p.D + 4 * 2 * kMaxT + 4 * kMaxT) * sizeof(float);
    if ((rc = launch_prologue_fast(vw, p, f, pw, w, smem, p.nblk1, fused_intra ? 1 : 0, st))) return rc;
  } else {
    if (pw.world > 1 || p.local_rows) return FOCAL_ESHAPE;      // the generic row kernels have no peer path
    const size_t smem = (size_t)kRowsPerBlock * p.nT * p.D * sizeof(float);
    static size_t configured = 0;
    if (smem > 48 * 1024 && smem > configured) {
      if (cudaFuncSetAttribute(prologue_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
        return cuda_ok("cudaFuncSetAttribute(prologue_kernel)");
      configured = smem;
    }
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
    if (vw) {
      if ((rc = launch_finalize_fast(vw, p, f, g, w, (rows + 3) / 4, st))) return rc;
    } else {
      const size_t smem = (size_t)kRowsPerBlock * (2 * p.nT + 1) * p.D * sizeof(float);
      static size_t configured = 0;
      if (smem > 48 * 1024 && smem > configured) {
        if (cudaFuncSetAttribute(finalize_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
          return cuda_ok("cudaFuncSetAttribute(finalize_kernel)");
        configured = smem;
      }
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
      return "unsupported shape: need B % S == 0, S a power of two <= 32, 2 <= D <= 512, 1 <= M <= 4, "
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
  nce_lse_kernel<<<lse_blocks(p, all_rows), 256, 0, static_cast<cudaStream_t>(stream)>>>(p, solo(ws), static_cast<uint8_t*>(ws), all_rows);
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
  if (!loss5) return FOCAL_EINVAL;
  return do_finalize(p, cfg->no_private, feats, grads, solo(ws), static_cast<uint8_t*>(ws), loss5,
                     static_cast<cudaStream_t>(stream));
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
  if (!peers || !loss5 || !p.local_rows) return FOCAL_EINVAL;
  if (peers->world < 1 || peers->world > kMaxPeers || peers->rank < 0 || peers->rank >= peers->world) return FOCAL_EINVAL;
  PeerWs pw{};
  pw.rank = peers->rank; pw.world = peers->world;
  for (int r = 0; r < pw.world; ++r) {
    if ((rc = check_ws(p, peers->ws[r], ws_bytes))) return rc;
    pw.ws[r] = static_cast<uint8_t*>(peers->ws[r]);
  }
  uint8_t* w = pw.ws[pw.rank];
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  FeatPtrs f;
  if ((rc = fill_feats(p, feats, f))) return rc;
  // phase 1: operands of the owned rows -> every workspace; the last block of the prologue announces epoch 1 and the
  // first Gram launch waits for every rank's announcement before it reads operands.  (No zero_pad launch: workspaces
  // from focal_b200_peer_alloc start zeroed and nobody ever writes a padding row.)
  if ((rc = do_prologue(p, cfg->no_private, f, pw, w, st, /*zero_pads=*/false))) return rc;
  // several ranks on one device: the waits become one-block launches of their own (see peer_wait_kernel)
  const bool split_wait = pw.world > 1 && ranks_share_a_device(peers);
  const int nwait = (pw.world > 1 && !split_wait) ? pw.world : 0;
  auto wait_launch = [&]() -> int {
    if (!split_wait) return FOCAL_OK;
    peer_wait_kernel<<<1, 32, 0, st>>>(p, pw);
    return cuda_ok("peer_wait_kernel");
  };
  const bool nce = (p.terms & FOCAL_TERM_NCE) != 0;
  const bool tmp = (p.terms & FOCAL_TERM_TEMPORAL) && !temporal_degenerate(p);
  // phase 2: row sums of the owned rows -> every workspace (announced by the last block of nce_lse).  The temporal launch
  // needs nothing from the peers beyond phase 1, so it runs between those stores and the launch that waits for them.
  if (nce || tmp) {
    if ((rc = wait_launch())) return rc;
  }
  if (nce) {
    if ((rc = launch_nce<NCE_FWD>(p, w, st, nwait))) return rc;
    nce_lse_kernel<<<lse_blocks(p, 0), 256, 0, st>>>(p, pw, w, 0);
    if ((rc = cuda_ok("nce_lse_kernel"))) return rc;
  }
  if (tmp) {
    const int wt = nce ? 0 : nwait;
    rc = p.need_grad ? launch_temporal<TMP_BWD>(p, w, st, p.grid_tmp, wt) : launch_temporal<TMP_FWD>(p, w, st, p.grid_tmp, wt);
    if (rc) return rc;
  }
  if (nce && p.need_grad) {
    if ((rc = wait_launch())) return rc;
    if ((rc = launch_nce<NCE_BWD>(p, w, st, nwait))) return rc;
  }
  // phase 3: gradients of the owned rows; loss partials all-reduced inside loss_reduce_kernel (third barrier)
  return do_finalize(p, cfg->no_private, feats, grads, pw, w, loss5, st);
}

}  // extern "C"

// ---------------------------------------------------------------------------------------------------------
// bring-up probe
// ---------------------------------------------------------------------------------------------------------
namespace {
__global__ void __launch_bounds__(128, 1) umma_probe_kernel(const uint8_t* a_img, uint32_t a_bytes, const uint8_t* b_img,
                                                            uint32_t b_bytes, uint32_t idesc, uint32_t a_lbo,
                                                            uint32_t a_sbo, uint32_t a_kstep, uint32_t b_lbo,
                                                            uint32_t b_sbo, uint32_t b_kstep, uint32_t ksteps,
                                                            uint32_t ncols, uint32_t a_via_st, float* d_out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar_load, bar_mma;
  __shared__ uint32_t tmem_slot;
  uint8_t* sa = smem;
  uint8_t* sb = smem + ((a_bytes + 1023) & ~1023u);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(&bar_load, 1);
    mbar_init(&bar_mma, 1);
    fence_mbar_init();
  }
  if (warp == 0) {
    tmem_alloc(&tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  const uint32_t a_tmem_col = 256;                 // A operand columns when it lives in tensor memory
  const bool tf32 = (a_via_st & 0x100u) != 0;      // kind::tf32 instead of kind::f16 (32-bit elements)
  a_via_st &= 0xffu;
  if (a_via_st == 1) {
    for (uint32_t o = threadIdx.x * 16; o < a_bytes; o += blockDim.x * 16)
      *reinterpret_cast<uint4*>(sa + o) = *reinterpret_cast<const uint4*>(a_img + o);
    fence_proxy_async_smem();
  } else if (a_via_st == 2) {
    const uint32_t W32 = a_bytes / 512;             // 32-bit words per row (2 bf16 or 1 tf32 element each)
    const uint32_t* rowp = reinterpret_cast<const uint32_t*>(a_img) + (size_t)threadIdx.x * W32;
    for (uint32_t g = 0; g < W32 / 32; ++g) {
      uint32_t r[32];
      for (int j = 0; j < 32; ++j) r[j] = rowp[g * 32 + j];
      tmem_st32(tmem + ((uint32_t)(warp * 32) << 16) + a_tmem_col + g * 32, r);
    }
    tmem_st_wait();
    tc_fence_before();
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    mbar_arrive_expect_tx(&bar_load, (a_via_st ? 0 : a_bytes) + b_bytes);
    if (!a_via_st) tma_load_1d(sa, a_img, a_bytes, &bar_load);
    tma_load_1d(sb, b_img, b_bytes, &bar_load);
    mbar_wait(&bar_load, 0);
    tc_fence_after();
    for (uint32_t k = 0; k < ksteps; ++k) {
      const uint64_t db = umma_smem_desc(smem_u32(sb) + k * b_kstep, b_lbo, b_sbo);
      if (tf32) {
        if (a_via_st == 2) umma_tf32_ts(tmem, tmem + a_tmem_col + k * 8, db, idesc, k > 0);
        else umma_tf32(tmem, umma_smem_desc(smem_u32(sa) + k * a_kstep, a_lbo, a_sbo), db, idesc, k > 0);
      } else if (a_via_st == 2) umma_bf16_ts(tmem, tmem + a_tmem_col + k * 8, db, idesc, k > 0);
      else umma_bf16(tmem, umma_smem_desc(smem_u32(sa) + k * a_kstep, a_lbo, a_sbo), db, idesc, k > 0);
    }
    umma_commit(&bar_mma);
  }
  mbar_wait(&bar_mma, 0);
  tc_fence_after();
  const int row = warp * 32 + lane;
  for (uint32_t c = 0; c < ncols; c += 32) {
    float v[32];
    tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + c, v);
    tmem_ld_wait();
    for (int j = 0; j < 32; ++j) d_out[(size_t)row * ncols + c + j] = v[j];
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}
}  // namespace

extern "C" int focal_b200_debug_umma(const void* a_img, uint32_t a_bytes, const void* b_img, uint32_t b_bytes,
                                     uint32_t idesc, uint32_t a_lbo, uint32_t a_sbo, uint32_t a_kstep_bytes,
                                     uint32_t b_lbo, uint32_t b_sbo, uint32_t b_kstep_bytes, uint32_t ksteps,
                                     uint32_t ncols, uint32_t a_via_st, float* d_out, void* stream) {
  if (!a_img || !b_img || !d_out || (a_bytes & 15) || (b_bytes & 15) || ncols % 32 || ncols > 256 || ncols == 0)
    return FOCAL_EINVAL;
  const uint32_t smem = ((a_bytes + 1023) & ~1023u) + b_bytes + 1024;
  if (smem > 220 * 1024) return FOCAL_EINVAL;
  if (cudaFuncSetAttribute(umma_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
    return cuda_ok("cudaFuncSetAttribute(umma_probe_kernel)");
  umma_probe_kernel<<<1, 128, smem, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const uint8_t*>(a_img), a_bytes, static_cast<const uint8_t*>(b_img), b_bytes, idesc, a_lbo, a_sbo,
      a_kstep_bytes, b_lbo, b_sbo, b_kstep_bytes, ksteps, ncols, a_via_st, d_out);
  return cuda_ok("umma_probe_kernel");
}

// ---------------------------------------------------------------------------------------------------------
// bring-up micro-benchmark: cycles per tcgen05.mma (M = 128) for a given N / operand placement.
// The issue loop is fully unrolled with precomputed descriptors so that the tensor pipe, not the issuing
// thread, is what is measured.
// ---------------------------------------------------------------------------------------------------------
namespace {
template <int N, int B_MN, int A_TMEM>
__global__ void __launch_bounds__(128, 1) umma_rate_kernel(uint32_t iters, long long* cycles, uint32_t sync_mode = 0) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5;
  for (uint32_t o = threadIdx.x * 16; o < 64 * 1024 + 128 * 1024; o += blockDim.x * 16)
    *reinterpret_cast<uint4*>(smem + o) = make_uint4(0, 0, 0, 0);
  fence_proxy_async_smem();
  __shared__ uint64_t bar_done[2], bar_ready;
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1); mbar_init(&bar_done[0], 1); mbar_init(&bar_done[1], 1); mbar_init(&bar_ready, 1);
    fence_mbar_init();
  }
  if (warp == 0) { tmem_alloc(&tmem_slot, 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  // CUTLASS-style issue: the whole warp runs the (warp-uniform) loop and one elected lane issues.  Every operand of
  // tcgen05.mma is derived from values the compiler can prove warp-uniform (shfl broadcast, kernel parameters,
  // shared-memory base), otherwise ptxas wraps each instruction in an ELECT / R2UR.BROADCAST / BRA.U.ANY waterfall.
  const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem, 0);
  if (__shfl_sync(0xffffffffu, warp, 0) == 0) {
    const uint32_t a_addr = smem_u32(smem), b_addr = smem_u32(smem + 64 * 1024);
    constexpr uint32_t idesc = umma_idesc(UMMA_BF16, 128, N, 0, B_MN);
    const uint64_t da0 = umma_smem_desc(a_addr, 16, 1024);
    const uint64_t db0 = B_MN ? umma_smem_desc(b_addr, 128 * 128, 1024) : umma_smem_desc(b_addr, 16, 1024);
    const long long t0 = clock64();
    for (uint32_t it = 0; it < iters; ++it) {
      if (sync_mode >= 1 && it >= 2) mbar_wait(&bar_done[it & 1], ((it >> 1) - 1) & 1);   // like s_empty / w_full
      if (sync_mode >= 2) tc_fence_after();
      const uint32_t d = tmem_u + (it & 1) * (A_TMEM ? 128 : 256) * (N > 128 && A_TMEM ? 0 : 1);
      if (elect_one()) {
#pragma unroll
        for (uint32_t k = 0; k < 8; ++k) {
          const uint64_t da = da0 + (((k >> 2) * 16384 + (k & 3) * 32) >> 4);
          const uint64_t db = db0 + ((B_MN ? k * 2048 : (k >> 2) * (N * 128) + (k & 3) * 32) >> 4);
          if (A_TMEM) umma_bf16_ts(d, tmem_u + 384 + k * 8, db, idesc, k > 0);
          else umma_bf16(d, da, db, idesc, k > 0);
        }
        if (sync_mode >= 1) umma_commit(&bar_done[it & 1]);                                // like s_full
        if (sync_mode >= 3) umma_commit(&bar_ready);                                       // like b_empty (never waited)
      }
      __syncwarp();
    }
    if (elect_one()) umma_commit(&bar);
    __syncwarp();
    mbar_wait(&bar, 0);
    const long long t1 = clock64();
    if ((threadIdx.x & 31) == 0) cycles[blockIdx.x] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}
template <int N, int B_MN, int A_TMEM>
int launch_rate(uint32_t iters, uint32_t grid, long long* cycles, cudaStream_t st, uint32_t sync_mode) {
  const uint32_t smem = 64 * 1024 + 128 * 1024 + 1024;
  auto k = umma_rate_kernel<N, B_MN, A_TMEM>;
  if (cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
    return cuda_ok("cudaFuncSetAttribute(umma_rate_kernel)");
  k<<<grid, 128, smem, st>>>(iters, cycles, sync_mode);
  return cuda_ok("umma_rate_kernel");
}
}  // namespace

// 8 MMAs (K = 16 each) per iteration; returns per-CTA cycles for `iters` iterations.
extern "C" int focal_b200_debug_umma_rate(uint32_t N, uint32_t flags, uint32_t iters, uint32_t sync_mode, uint32_t grid,
                                          long long* cycles, void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const uint32_t b_mn = flags & 1, a_tmem = (flags >> 1) & 1;
#define FB_RATE(NN)                                                                      \
  if (N == NN) {                                                                         \
    if (!b_mn && !a_tmem) return launch_rate<NN, 0, 0>(iters, grid, cycles, st, sync_mode);         \
    if (b_mn && !a_tmem) return launch_rate<NN, 1, 0>(iters, grid, cycles, st, sync_mode);          \
    if (!b_mn && a_tmem) return launch_rate<NN, 0, 1>(iters, grid, cycles, st, sync_mode);          \
    return launch_rate<NN, 1, 1>(iters, grid, cycles, st, sync_mode);                               \
  }
  FB_RATE(32) FB_RATE(64) FB_RATE(96) FB_RATE(128) FB_RATE(192) FB_RATE(256)
#undef FB_RATE
  return FOCAL_EINVAL;
}

// ---------------------------------------------------------------------------------------------------------
// bring-up micro-benchmark: L2 -> shared-memory throughput of linear bulk copies (cp.async.bulk) per SM
// ---------------------------------------------------------------------------------------------------------
namespace {
__global__ void __launch_bounds__(64, 1) tma_rate_kernel(const uint8_t* src, uint32_t span_bytes, uint32_t copy_bytes,
                                                         uint32_t copies_per_stage, uint32_t stages, uint32_t iters,
                                                         long long* cycles) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t full[8], empty[8];
  if (threadIdx.x == 0) {
    for (int i = 0; i < 8; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    fence_mbar_init();
  }
  __syncthreads();
  const uint32_t stage_bytes = copy_bytes * copies_per_stage;
  const int warp_u = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  if (warp_u == 0) {                              // producer: warp-uniform loop, one elected lane issues
    const long long t0 = clock64();
    uint32_t off = blockIdx.x * 65536u;
    for (uint32_t n = 0; n < iters; ++n) {
      const uint32_t st = n % stages;
      mbar_wait(&empty[st], ((n / stages) & 1) ^ 1);
      if (elect_one()) {
        mbar_arrive_expect_tx(&full[st], stage_bytes);
        for (uint32_t c = 0; c < copies_per_stage; ++c) {
          tma_load_1d(smem + st * stage_bytes + c * copy_bytes, src + (off & (span_bytes - 1)), copy_bytes, &full[st]);
          off += copy_bytes;
        }
      }
      off = __shfl_sync(0xffffffffu, off, 0);
    }
    if ((threadIdx.x & 31) == 0) cycles[blockIdx.x * 2] = clock64() - t0;
  } else if (threadIdx.x == 32) {                 // consumer: frees the stage as soon as the bytes have landed
    const long long t0 = clock64();
    for (uint32_t n = 0; n < iters; ++n) {
      const uint32_t st = n % stages;
      mbar_wait(&full[st], (n / stages) & 1);
      mbar_arrive(&empty[st]);
    }
    cycles[blockIdx.x * 2 + 1] = clock64() - t0;
  }
}
}  // namespace

extern "C" int focal_b200_debug_tma_rate(const void* src, uint32_t span_bytes, uint32_t copy_bytes,
                                         uint32_t copies_per_stage, uint32_t stages, uint32_t iters, uint32_t grid,
                                         long long* cycles, void* stream) {
  const uint32_t smem = copy_bytes * copies_per_stage * stages + 1024;
  if (smem > 220 * 1024 || stages > 8) return FOCAL_EINVAL;
  if (cudaFuncSetAttribute(tma_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
    return cuda_ok("cudaFuncSetAttribute(tma_rate_kernel)");
  tma_rate_kernel<<<grid, 64, smem, static_cast<cudaStream_t>(stream)>>>(static_cast<const uint8_t*>(src), span_bytes,
                                                                        copy_bytes, copies_per_stage, stages, iters,
                                                                        cycles);
  return cuda_ok("tma_rate_kernel");
}
