// Device-side synchronisation between the ranks of a row-sharded job (flags in NVLink peer memory).
#pragma once
#include <cstdio>

#include "plan.h"

namespace fb {

// Device-side barrier over the ranks of a row-sharded job, split in two phases so that no launch exists only to
// synchronise.  Flags live in every rank's workspace: slot r of rank q's array = the last epoch rank r announced to q.
//   announce: store the rank's epoch into its slot of every peer's array (st.release.sys after a system fence:
//             everything this rank wrote to peer memory before is visible first) -- done by the launch AFTER the one
//             that produced the data (peer_epoch_bump / peer_announce_epoch);
//   wait:     spin (ld.acquire.sys, bounded: a lost rank traps instead of hanging) until every peer's slot in the own
//             array has reached the own epoch.
// Epochs are counted on the device (slot 8), so the sequence replays unchanged inside a CUDA graph.
__device__ __forceinline__ void peer_announce(const Plan& p, const PeerWs& pw, uint32_t e) {     // threads < world
  if ((int)threadIdx.x < pw.world) {
    __threadfence_system();
    uint32_t* dst = reinterpret_cast<uint32_t*>(pw.ws[threadIdx.x] + p.bar_off) + pw.rank;
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(dst), "r"(e) : "memory");
  }
}
__device__ __forceinline__ void peer_wait(const uint8_t* ws, const Plan& p, int world, int rank) {   // threads < world
  if ((int)threadIdx.x < world) {
    const uint32_t* mine = reinterpret_cast<const uint32_t*>(ws + p.bar_off);
    uint32_t e;
    asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(e) : "l"(mine + 8) : "memory");
    const uint32_t* src = mine + threadIdx.x;
    uint64_t t0 = 0;
    for (uint32_t spins = 0;; ++spins) {
      uint32_t v;
      asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(src) : "memory");
      if ((int32_t)(v - e) >= 0) break;
      if ((spins & 1023) == 1023) {
        uint64_t now;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
        if (t0 == 0) t0 = now;
        else if (now - t0 > 20000000000ull) {          // 20 s: a rank died or fell out of step
          printf("focal_b200: peer barrier timed out (rank %d waiting for rank %d, epoch %u, saw %u)\n", rank,
                 (int)threadIdx.x, e, v);
          __trap();
        }
      }
    }
  }
}
// Full barrier by ONE block of >= kMaxPeers threads (all ranks in step).
__device__ __forceinline__ void peer_barrier(const Plan& p, const PeerWs& pw) {
  __shared__ uint32_t epoch_sh;
  uint32_t* mine = reinterpret_cast<uint32_t*>(pw.ws[pw.rank] + p.bar_off);
  __syncthreads();
  if (threadIdx.x == 0) { epoch_sh = mine[8] + 1; mine[8] = epoch_sh; __threadfence(); }
  __syncthreads();
  peer_announce(p, pw, epoch_sh);
  peer_wait(pw.ws[pw.rank], p, pw.world, pw.rank);
  __syncthreads();
}
// Producer launches (prologue, nce_lse) only COUNT the epoch: one thread of the launch bumps the rank's counter; the
// announcement itself is sent by block 0 of the NEXT launch on the stream (peer_announce_epoch below), when stream order
// has already completed every store of the producer.  (The first version announced from the last block to finish, which
// needed a system fence + an atomic at the end of EVERY block: measured 11-15 us of the prologue and 6 us of nce_lse at
// 4 GPUs -- profiles/r2_prologue_dbg_4.txt.)
__device__ __forceinline__ void peer_epoch_bump(const Plan& p, uint8_t* ws) {
  if (blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) {
    uint32_t* mine = reinterpret_cast<uint32_t*>(ws + p.bar_off);
    mine[8] = mine[8] + 1;
  }
}
// Block 0 of the launch that follows a producer: send the rank's current epoch to every peer.  `peer_ws` = the R mapped
// workspaces, own rank included.
__device__ __forceinline__ void peer_announce_epoch(const Plan& p, uint8_t* const* peer_ws, int world, int rank) {   // threads < world
  if ((int)threadIdx.x < world) {
    uint32_t e;
    asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(e) : "l"(reinterpret_cast<const uint32_t*>(peer_ws[rank] + p.bar_off) + 8) : "memory");
    __threadfence_system();
    uint32_t* dst = reinterpret_cast<uint32_t*>(peer_ws[threadIdx.x] + p.bar_off) + rank;
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(dst), "r"(e) : "memory");
  }
}

}  // namespace fb
