// C ABI of the B200-native FOCAL loss hot path (see include/focal_b200.h).
#include <cstdio>
#include <cuda_runtime.h>

#include "../../include/focal_b200.h"
#include "gram_kernel.cuh"
#include "plan.h"
#include "row_kernels.cuh"

using namespace fb;

namespace {

int device_sms() {
  int dev = 0, sms = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 148;
  if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) return 148;
  return sms;
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

template <int MODE, int BN, int KB, int SEQ>
int launch_gram(const Plan& p, uint8_t* ws, cudaStream_t st, int n_items) {
  using L = GramSmem<BN, KB>;
  auto kfn = gram_kernel<MODE, BN, KB, SEQ>;
  static bool configured = false;     // per instantiation; the attribute is sticky per context
  if (!configured) {
    if (cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L::kDynamic) != cudaSuccess)
      return cuda_ok("cudaFuncSetAttribute(gram_kernel)");
    configured = true;
  }
  if (n_items <= 0) return FOCAL_OK;
  const int grid = n_items < p.num_sms ? n_items : p.num_sms;     // persistent: one CTA per SM
  kfn<<<grid, kGramThreads, L::kDynamic, st>>>(p, ws);
  return cuda_ok("gram_kernel launch");
}

template <int MODE, int SEQ>
int launch_gram_kb(const Plan& p, uint8_t* ws, cudaStream_t st, int kb, int n_items) {
  switch (kb) {
    case 1: return launch_gram<MODE, 128, 1, SEQ>(p, ws, st, n_items);
    case 2: return launch_gram<MODE, 128, 2, SEQ>(p, ws, st, n_items);
    case 3: return launch_gram<MODE, 64, 3, SEQ>(p, ws, st, n_items);
    case 4: return launch_gram<MODE, 64, 4, SEQ>(p, ws, st, n_items);
  }
  return FOCAL_ESHAPE;
}

template <int MODE>
int launch_temporal(const Plan& p, uint8_t* ws, cudaStream_t st, int n_items) {
  switch (p.S) {
    case 2: return launch_gram_kb<MODE, 2>(p, ws, st, p.kbFull, n_items);
    case 4: return launch_gram_kb<MODE, 4>(p, ws, st, p.kbFull, n_items);
    case 8: return launch_gram_kb<MODE, 8>(p, ws, st, p.kbFull, n_items);
    case 16: return launch_gram_kb<MODE, 16>(p, ws, st, p.kbFull, n_items);
    case 32: return launch_gram_kb<MODE, 32>(p, ws, st, p.kbFull, n_items);
  }
  return FOCAL_ESHAPE;
}

int nce_items(const Plan& p, bool fwd) {
  const int t0 = p.seq0 / kTileM, t1 = (p.seq1 + kTileM - 1) / kTileM;
  return p.nProb * p.S * 2 * (t1 - t0) * (fwd ? p.nsplit_fwd : 1);
}
int tmp_items(const Plan& p) {
  const int t0 = (p.seq0 * p.S) / kTileM, t1 = (p.seq1 * p.S + kTileM - 1) / kTileM;
  return p.nT * (t1 - t0);
}

// InfoNCE launches are grouped by operand width (all problems share it unless noPrivate mixes D and D/2)
int nce_kb(const Plan& p) { return p.ops[p.probs[0].opA].kb; }
bool nce_uniform(const Plan& p) {
  for (int q = 1; q < p.nProb; ++q)
    if (p.ops[p.probs[q].opA].kb != nce_kb(p)) return false;
  return true;
}

int fill_feats(const Plan& p, const float* const* feats, FeatPtrs& f) {
  if (!feats) return FOCAL_EINVAL;
  for (int t = 0; t < p.nT; ++t) {
    if (!feats[t] || (reinterpret_cast<uintptr_t>(feats[t]) & 15)) return FOCAL_EINVAL;
    f.x[t] = feats[t];
  }
  for (int t = p.nT; t < kMaxT; ++t) f.x[t] = nullptr;
  return FOCAL_OK;
}

}  // namespace

extern "C" {

int focal_b200_abi_version(void) { return FOCAL_B200_ABI_VERSION; }

const char* focal_b200_strerror(int code) {
  switch (code) {
    case FOCAL_OK: return "ok";
    case FOCAL_EINVAL: return "invalid argument";
    case FOCAL_ESHAPE:
      return "unsupported shape: need B % S == 0, S a power of two <= 32, 2 <= D <= 256, 1 <= M <= 4, "
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
  return FOCAL_OK;
}

int focal_b200_prologue(const FocalCfg* cfg, const float* const* feats, void* ws, size_t ws_bytes, void* stream) {
  Plan p;
  int rc = make_plan(cfg, p);
  if (rc) return rc;
  if ((rc = check_ws(p, ws, ws_bytes))) return rc;
  FeatPtrs f;
  if ((rc = fill_feats(p, feats, f))) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  uint8_t* w = static_cast<uint8_t*>(ws);
  if (p.bpad != p.b || p.Bpad != p.B) {
    zero_pad_kernel<<<64, 256, 0, st>>>(p, w);
    if ((rc = cuda_ok("zero_pad_kernel"))) return rc;
  }
  const size_t smem = (size_t)kRowsPerBlock * p.nT * p.D * sizeof(float);
  static size_t configured = 0;
  if (smem > 48 * 1024 && smem > configured) {
    if (cudaFuncSetAttribute(prologue_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
      return cuda_ok("cudaFuncSetAttribute(prologue_kernel)");
    configured = smem;
  }
  prologue_kernel<<<p.nblk1, 32 * kRowsPerBlock, smem, st>>>(p, f, w);
  if ((rc = cuda_ok("prologue_kernel"))) return rc;
  if ((p.terms & FOCAL_TERM_TEMPORAL) && !temporal_degenerate(p)) {
    const long warps = (long)p.nT * p.b;
    intra_kernel<<<(unsigned)((warps + 3) / 4), 128, 0, st>>>(p, f, w);
    if ((rc = cuda_ok("intra_kernel"))) return rc;
  }
  return FOCAL_OK;
}

int focal_b200_nce_rowsum(const FocalCfg* cfg, void* ws, size_t ws_bytes, void* stream) {
  Plan p;
  int rc = make_plan(cfg, p);
  if (rc) return rc;
  if ((rc = check_ws(p, ws, ws_bytes))) return rc;
  if (!(p.terms & FOCAL_TERM_NCE)) return FOCAL_OK;
  if (!nce_uniform(p)) return FOCAL_ESHAPE;
  return launch_gram_kb<NCE_FWD, 0>(p, static_cast<uint8_t*>(ws), static_cast<cudaStream_t>(stream), nce_kb(p),
                                    nce_items(p, true));
}

int focal_b200_nce_lse(const FocalCfg* cfg, void* ws, size_t ws_bytes, int all_rows, void* stream) {
  Plan p;
  int rc = make_plan(cfg, p);
  if (rc) return rc;
  if ((rc = check_ws(p, ws, ws_bytes))) return rc;
  if (!(p.terms & FOCAL_TERM_NCE)) return FOCAL_OK;
  nce_lse_kernel<<<p.nblk2, 256, 0, static_cast<cudaStream_t>(stream)>>>(p, static_cast<uint8_t*>(ws), all_rows);
  return cuda_ok("nce_lse_kernel");
}

int focal_b200_nce_grad(const FocalCfg* cfg, void* ws, size_t ws_bytes, void* stream) {
  Plan p;
  int rc = make_plan(cfg, p);
  if (rc) return rc;
  if ((rc = check_ws(p, ws, ws_bytes))) return rc;
  if (!(p.terms & FOCAL_TERM_NCE) || !p.need_grad) return FOCAL_OK;
  if (!nce_uniform(p)) return FOCAL_ESHAPE;
  return launch_gram_kb<NCE_BWD, 0>(p, static_cast<uint8_t*>(ws), static_cast<cudaStream_t>(stream), nce_kb(p),
                                    nce_items(p, false));
}

int focal_b200_temporal(const FocalCfg* cfg, void* ws, size_t ws_bytes, void* stream) {
  Plan p;
  int rc = make_plan(cfg, p);
  if (rc) return rc;
  if ((rc = check_ws(p, ws, ws_bytes))) return rc;
  if (!(p.terms & FOCAL_TERM_TEMPORAL) || temporal_degenerate(p)) return FOCAL_OK;
  uint8_t* w = static_cast<uint8_t*>(ws);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  return p.need_grad ? launch_temporal<TMP_BWD>(p, w, st, tmp_items(p)) : launch_temporal<TMP_FWD>(p, w, st, tmp_items(p));
}

int focal_b200_finalize(const FocalCfg* cfg, const float* const* feats, void* ws, size_t ws_bytes, float* loss5,
                        float* const* grads, void* stream) {
  Plan p;
  int rc = make_plan(cfg, p);
  if (rc) return rc;
  if ((rc = check_ws(p, ws, ws_bytes))) return rc;
  if (!loss5) return FOCAL_EINVAL;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  uint8_t* w = static_cast<uint8_t*>(ws);
  if (p.need_grad) {
    FeatPtrs f;
    if ((rc = fill_feats(p, feats, f))) return rc;
    if (!grads) return FOCAL_EINVAL;
    GradPtrs g;
    for (int t = 0; t < kMaxT; ++t) g.g[t] = nullptr;
    for (int t = 0; t < p.nT; ++t) {
      if (!grads[t] || (reinterpret_cast<uintptr_t>(grads[t]) & 15)) return FOCAL_EINVAL;
      g.g[t] = grads[t];
    }
    const size_t smem = (size_t)kRowsPerBlock * (2 * p.nT + 1) * p.D * sizeof(float);
    static size_t configured = 0;
    if (smem > 48 * 1024 && smem > configured) {
      if (cudaFuncSetAttribute(finalize_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
        return cuda_ok("cudaFuncSetAttribute(finalize_kernel)");
      configured = smem;
    }
    const int rows = (p.seq1 - p.seq0) * p.S;
    finalize_kernel<<<(rows + kRowsPerBlock - 1) / kRowsPerBlock, 32 * kRowsPerBlock, smem, st>>>(p, f, g, w);
    if ((rc = cuda_ok("finalize_kernel"))) return rc;
  }
  loss_reduce_kernel<<<1, 256, 0, st>>>(p, w, loss5, p.nblk2, temporal_degenerate(p) ? 1 : 0);
  return cuda_ok("loss_reduce_kernel");
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
  if (a_via_st) {
    for (uint32_t o = threadIdx.x * 16; o < a_bytes; o += blockDim.x * 16)
      *reinterpret_cast<uint4*>(sa + o) = *reinterpret_cast<const uint4*>(a_img + o);
    fence_proxy_async_smem();
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    mbar_arrive_expect_tx(&bar_load, (a_via_st ? 0 : a_bytes) + b_bytes);
    if (!a_via_st) tma_load_1d(sa, a_img, a_bytes, &bar_load);
    tma_load_1d(sb, b_img, b_bytes, &bar_load);
    mbar_wait(&bar_load, 0);
    tc_fence_after();
    for (uint32_t k = 0; k < ksteps; ++k)
      umma_bf16(tmem, umma_smem_desc(smem_u32(sa) + k * a_kstep, a_lbo, a_sbo),
                umma_smem_desc(smem_u32(sb) + k * b_kstep, b_lbo, b_sbo), idesc, k > 0);
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
