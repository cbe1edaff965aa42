/* Bring-up entry points (libfocal_bringup.so; not part of the product ABI in include/focal_b200.h). */
#ifndef FOCAL_B200_BRINGUP_H_
#define FOCAL_B200_BRINGUP_H_
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* Runs `ksteps` tcgen05.mma (M=128) on caller-provided shared-memory images and returns the 128 x ncols fp32
 * accumulator.  flags: bits [0,8) A placement (0 = TMA copy, 1 = st.shared, 2 = tensor memory), 0x100 = kind::tf32,
 * bits [12,15) / [16,19) = shared-memory descriptor layout type + 1 of A / B (0 = SWIZZLE_128B). */
int focal_b200_debug_umma(const void* a_img, uint32_t a_bytes, const void* b_img, uint32_t b_bytes, uint32_t idesc,
                          uint32_t a_lbo, uint32_t a_sbo, uint32_t a_kstep_bytes, uint32_t b_lbo, uint32_t b_sbo,
                          uint32_t b_kstep_bytes, uint32_t ksteps, uint32_t ncols, uint32_t flags, float* d_out,
                          void* stream);
/* cycles per tcgen05.mma for a given N / operand placement / per-tile barrier traffic */
int focal_b200_debug_umma_rate(uint32_t N, uint32_t flags, uint32_t iters, uint32_t sync_mode, uint32_t grid,
                               long long* cycles, void* stream);
/* tensor-pipe time of one column tile of the fused temporal kernel (16 SS MMAs of N = BN, then BN/16 TS MMAs of N = 256),
 * no epilogue; mode 0 = alternating as the kernel issues them, 1 = only #1, 2 = only #2, 3 = two tiles batched */
int focal_b200_debug_umma_tile_rate(uint32_t BN, uint32_t mode, uint32_t iters, uint32_t grid, long long* cycles,
                                    void* stream);
/* per-SM throughput of linear TMA bulk copies */
int focal_b200_debug_tma_rate(const void* src, uint32_t span_bytes, uint32_t copy_bytes, uint32_t copies_per_stage,
                              uint32_t stages, uint32_t iters, uint32_t grid, long long* cycles, void* stream);
/* SM stores of `bytes` (16 per thread and iteration) at offset off0 of n destination buffers (peer-mapped or multicast
 * addresses): mode 0 = chunk-major unicast, 1 = multimem.st to dsts[0], 2 = destination-major unicast */
int focal_b200_debug_peer_store(void* const* dsts, int n, uint32_t off0, uint32_t bytes, int mode, int grid,
                                void* stream);
/* enqueues `iters` launches of an (almost) empty 148 x 576 kernel: flags 1 = 10.7 KB by-value parameters, 2 = 200 KB
 * dynamic shared memory, 4 = tensor memory allocated + freed, 8 = a small-shared-memory 1024-block launch in between */
int focal_b200_debug_launch_floor(uint32_t flags, uint32_t iters, uint32_t* out, void* stream);

#ifdef __cplusplus
}
#endif
#endif
