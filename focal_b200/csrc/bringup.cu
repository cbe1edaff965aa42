// Bring-up probes and micro-benchmarks (NOT part of the product library): built into libfocal_bringup.so by
// `python -m focal_b200.build` and used only by tests/test_gpu_probe.py and tools/{umma_rate,tma_rate,tf32_probe}.py.
// They pin the tcgen05 descriptor encodings on real hardware and measure issue / copy rates.
#include <cstdio>
#include <cuda_runtime.h>

#include "bringup.h"
#include "ptx.cuh"

using namespace fb;

namespace {
int cuda_ok(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    fprintf(stderr, "focal_b200 bring-up: %s failed: %s\n", what, cudaGetErrorString(e));
    return -3;
  }
  return 0;
}
constexpr int FOCAL_EINVAL = -1;
}  // namespace

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
  // bits [12,15) / [16,19): smem descriptor layout type + 1 of A / B (0 = the default, SWIZZLE_128B)
  const uint32_t a_lay = ((a_via_st >> 12) & 7u) ? ((a_via_st >> 12) & 7u) - 1 : (uint32_t)UMMA_LAYOUT_SW128;
  const uint32_t b_lay = ((a_via_st >> 16) & 7u) ? ((a_via_st >> 16) & 7u) - 1 : (uint32_t)UMMA_LAYOUT_SW128;
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
      const uint64_t db = umma_smem_desc(smem_u32(sb) + k * b_kstep, b_lbo, b_sbo, b_lay);
      if (tf32) {
        if (a_via_st == 2) umma_tf32_ts(tmem, tmem + a_tmem_col + k * 8, db, idesc, k > 0);
        else umma_tf32(tmem, umma_smem_desc(smem_u32(sa) + k * a_kstep, a_lbo, a_sbo, a_lay), db, idesc, k > 0);
      } else if (a_via_st == 2) umma_bf16_ts(tmem, tmem + a_tmem_col + k * 8, db, idesc, k > 0);
      else umma_bf16(tmem, umma_smem_desc(smem_u32(sa) + k * a_kstep, a_lbo, a_sbo, a_lay), db, idesc, k > 0);
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
// bring-up micro-benchmark: the tensor-pipe schedule of one column tile of the fused temporal kernel, without any
// epilogue: UMMA #1 = 16 x (SS, N = BN, K-major B) into S stage (tile & 1), UMMA #2 = BN/16 x (TS, N = 256, MN-major
// B, A from the S stage) into the O accumulator at column 0.  mode 0: #1 and #2 alternate tile by tile (what the
// kernel issues); 1: only #1; 2: only #2; 3: #1 of two tiles back to back, then #2 of both.  What the pipe itself
// needs per tile is the ceiling of the kernel's tensor-pipe share.
// ---------------------------------------------------------------------------------------------------------
namespace {
template <int BN>
__global__ void __launch_bounds__(128, 1) umma_tile_rate_kernel(uint32_t iters, uint32_t mode, long long* cycles) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5;
  for (uint32_t o = threadIdx.x * 16; o < 64 * 1024 + 64 * 1024; o += blockDim.x * 16)
    *reinterpret_cast<uint4*>(smem + o) = make_uint4(0, 0, 0, 0);
  fence_proxy_async_smem();
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  if (warp == 0) { tmem_alloc(&tmem_slot, 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_slot, 0);
  if (__shfl_sync(0xffffffffu, warp, 0) == 0) {
    const uint32_t a_addr = smem_u32(smem), b_addr = smem_u32(smem + 64 * 1024);
    constexpr uint32_t idesc1 = umma_idesc(UMMA_BF16, 128, BN, 0, 0);
    constexpr uint32_t idesc2 = umma_idesc(UMMA_BF16, 128, 256, 0, 1);
    const uint64_t da0 = umma_smem_desc(a_addr, 16, 1024);
    const uint64_t db0 = umma_smem_desc(b_addr, 16, 1024);
    const uint64_t dm0 = umma_smem_desc(b_addr, BN * 128, 1024);
    auto u1 = [&](uint32_t tile) {
      const uint32_t d = tmem_u + 256 + (tile & 1) * BN;
#pragma unroll
      for (uint32_t k = 0; k < 16; ++k)
        umma_bf16(d, da0 + (((k >> 2) * 16384 + (k & 3) * 32) >> 4), db0 + (((k >> 2) * (BN * 128) + (k & 3) * 32) >> 4),
                  idesc1, k > 0);
    };
    auto u2 = [&](uint32_t tile) {
      const uint32_t a = tmem_u + 256 + (tile & 1) * BN;
#pragma unroll
      for (uint32_t k = 0; k < BN / 16; ++k) umma_bf16_ts(tmem_u, a + k * 16, dm0 + ((k * 2048) >> 4), idesc2, 1u);
    };
    const long long t0 = clock64();
    for (uint32_t it = 0; it < iters; it += 2) {
      if (elect_one()) {
        if (mode == 0) { u1(it); u2(it); u1(it + 1); u2(it + 1); }
        else if (mode == 1) { u1(it); u1(it + 1); }
        else if (mode == 2) { u2(it); u2(it + 1); }
        else { u1(it); u1(it + 1); u2(it); u2(it + 1); }
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
  if (warp == 0) tmem_dealloc(tmem_u, 512);
}
}  // namespace

extern "C" int focal_b200_debug_umma_tile_rate(uint32_t BN, uint32_t mode, uint32_t iters, uint32_t grid, long long* cycles,
                                               void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const uint32_t smem = 64 * 1024 + 64 * 1024 + 1024;
#define FB_TILE_RATE(NN)                                                                                        \
  if (BN == NN) {                                                                                               \
    auto k = umma_tile_rate_kernel<NN>;                                                                         \
    if (cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)         \
      return cuda_ok("cudaFuncSetAttribute(umma_tile_rate_kernel)");                                            \
    k<<<grid, 128, smem, st>>>(iters & ~1u, mode, cycles);                                                      \
    return cuda_ok("umma_tile_rate_kernel");                                                                    \
  }
  FB_TILE_RATE(64) FB_TILE_RATE(80) FB_TILE_RATE(96) FB_TILE_RATE(128)
#undef FB_TILE_RATE
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

// ---------------------------------------------------------------------------------------------------------
// peer-memory store rate (tools/peer_store_rate.py): how fast can SM stores push one rank's operand slice into the
// workspaces of its peers -- one plain 16-byte store per destination, or one NVSwitch multicast store (multimem.st)
// ---------------------------------------------------------------------------------------------------------
namespace {
struct PeerDsts { uint8_t* p[8]; };
// mode 0: every 16-byte chunk stored to each of the n destinations (chunk-major, as the row kernels do);
// mode 1: multimem.st of every chunk to p[0] (a multicast address); mode 2: destination-major unicast.
__global__ void __launch_bounds__(256) peer_store_kernel(PeerDsts d, int n, uint32_t off0, uint32_t bytes, int mode) {
  const uint32_t stride = gridDim.x * blockDim.x * 16u;
  const uint4 v = make_uint4(threadIdx.x, blockIdx.x, 0x3f800000u, 0x3f800000u);
  if (mode == 1) {
    for (uint32_t o = (blockIdx.x * blockDim.x + threadIdx.x) * 16u; o < bytes; o += stride)
      asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(d.p[0] + off0 + o),
                   "f"(__uint_as_float(v.x)), "f"(__uint_as_float(v.y)), "f"(__uint_as_float(v.z)),
                   "f"(__uint_as_float(v.w)) : "memory");
  } else if (mode == 0) {
    for (uint32_t o = (blockIdx.x * blockDim.x + threadIdx.x) * 16u; o < bytes; o += stride)
      for (int k = 0; k < n; ++k) *reinterpret_cast<uint4*>(d.p[k] + off0 + o) = v;
  } else {
    for (int k = 0; k < n; ++k)
      for (uint32_t o = (blockIdx.x * blockDim.x + threadIdx.x) * 16u; o < bytes; o += stride)
        *reinterpret_cast<uint4*>(d.p[k] + off0 + o) = v;
  }
  __syncthreads();
  if (threadIdx.x == 0) __threadfence_system();
}
}  // namespace

extern "C" int focal_b200_debug_peer_store(void* const* dsts, int n, uint32_t off0, uint32_t bytes, int mode, int grid,
                                           void* stream) {
  if (n < 1 || n > 8 || grid < 1) return FOCAL_EINVAL;
  PeerDsts d{};
  for (int k = 0; k < n; ++k) d.p[k] = static_cast<uint8_t*>(dsts[k]);
  peer_store_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(d, n, off0, bytes, mode);
  return cuda_ok("peer_store_kernel");
}

// ---------------------------------------------------------------------------------------------------------
// launch floor (tools/launch_floor.py): what does a launch cost on the device, as a function of what the kernel asks
// for -- a 10 KB by-value parameter block, ~200 KB of dynamic shared memory, tensor memory -- and of what ran before it
// ---------------------------------------------------------------------------------------------------------
namespace {
struct BigParam { uint32_t w[2688]; };     // 10752 bytes: the size of Plan + ProbSel
__global__ void __launch_bounds__(576, 1) floor_small_kernel(uint32_t* out, uint32_t tm) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint32_t slot;
  if (tm) {
    if (threadIdx.x < 32) { tmem_alloc(&slot, 512); tmem_relinquish(); }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (threadIdx.x < 32) tmem_dealloc(slot, 512);
  }
  if (threadIdx.x == 0) { smem_raw[0] = 1; if (out && smem_raw[0] == 77) out[blockIdx.x] = 1; }
}
__global__ void __launch_bounds__(576, 1) floor_big_kernel(const __grid_constant__ BigParam bp, uint32_t* out, uint32_t tm) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint32_t slot;
  if (tm) {
    if (threadIdx.x < 32) { tmem_alloc(&slot, 512); tmem_relinquish(); }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (threadIdx.x < 32) tmem_dealloc(slot, 512);
  }
  if (threadIdx.x == 0) { smem_raw[0] = (uint8_t)bp.w[blockIdx.x]; if (out && smem_raw[0] == 77) out[blockIdx.x] = bp.w[2687]; }
}
__global__ void __launch_bounds__(256, 2) floor_row_kernel(uint32_t* out) {
  extern __shared__ uint8_t smem_raw[];
  if (threadIdx.x == 0) { smem_raw[0] = 1; if (out && smem_raw[0] == 77) out[blockIdx.x] = 1; }
}
}  // namespace

// Enqueues `iters` launches.  flags: 1 = 10.7 KB by-value parameter block, 2 = 200 KB dynamic shared memory, 4 = tensor
// memory allocated and freed, 8 = a 1024-block x 256-thread launch with 16 KB of shared memory between any two (the row
// kernels between the Gram launches: a different shared-memory carve-out).
extern "C" int focal_b200_debug_launch_floor(uint32_t flags, uint32_t iters, uint32_t* out, void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const size_t smem = (flags & 2) ? 200 * 1024 : 0;
  if (cudaFuncSetAttribute(floor_small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024) != cudaSuccess ||
      cudaFuncSetAttribute(floor_big_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024) != cudaSuccess)
    return cuda_ok("cudaFuncSetAttribute(floor kernels)");
  static BigParam bp{};
  for (uint32_t i = 0; i < iters; ++i) {
    if (flags & 1) floor_big_kernel<<<148, 576, smem, st>>>(bp, out, (flags & 4) ? 1u : 0u);
    else floor_small_kernel<<<148, 576, smem, st>>>(out, (flags & 4) ? 1u : 0u);
    if (flags & 8) floor_row_kernel<<<1024, 256, 16 * 1024, st>>>(out);
  }
  return cuda_ok("launch floor kernels");
}
