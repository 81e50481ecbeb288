// Slot / phase arithmetic of the cluster Gauss-Jordan inverse (linalg.cu: gj_inverse_kernel).  Pure integer logic,
// __host__ __device__: the kernel uses it for every mbarrier wait, tests/cpp/test_gj_phases.cu replays the whole protocol
// on the CPU for every N and checks it.
#pragma once

#include <cstdint>

namespace isr {

constexpr int kGjSlots = 4;   // pivot blocks in flight: the owner chain runs ahead of the remote consumers by up to 4 blocks

// Phase bookkeeping of the slot barriers for one CTA (blocks [Bs, Be) of NB are produced locally).  Every mbarrier wait
// needs the parity of the number of phases its barrier has completed before: these functions count the earlier uses.
// __host__ __device__: tests/cpp/test_gj_phases.cu replays the whole protocol on the CPU for every N and checks them.
struct GjPhases {
  int NB, Bs, Be;
  __host__ __device__ bool owned(int B) const { return B >= Bs && B < Be; }
  __host__ __device__ static int below(int x, int slot) { return x <= 0 ? 0 : (x + kGjSlots - 1 - slot) / kGjSlots; }   // #{y in [0, x): y = slot mod 4}
  // next block after B in the same slot that arrives from another CTA (its rfull phase needs an expect_tx), or -1
  __host__ __device__ int next_remote(int B) const {
    for (B += kGjSlots; B < NB; B += kGjSlots)
      if (!owned(B)) return B;
    return -1;
  }
  // owner of block B >= 4 waits for the credits of block B - 4: earlier waits on this CTA's empty[slot]
  __host__ __device__ uint32_t empty_parity(int B) const {
    const int q = B & (kGjSlots - 1);
    int B0 = Bs > kGjSlots ? Bs : kGjSlots;
    B0 += (q - B0) & (kGjSlots - 1);
    return (uint32_t)((B - B0) / kGjSlots) & 1u;
  }
  // consumer of a locally produced block: earlier uses of lfull[slot]
  __host__ __device__ uint32_t lfull_parity(int B) const {
    const int q = B & (kGjSlots - 1);
    return (uint32_t)(below(B, q) - below(Bs, q)) & 1u;
  }
  // consumer of a block from another CTA: earlier uses of rfull[slot]
  __host__ __device__ uint32_t rfull_parity(int B) const {
    const int q = B & (kGjSlots - 1);
    const int mine_before = B < Bs ? 0 : below(Be, q) - below(Bs, q);
    return (uint32_t)(below(B, q) - mine_before) & 1u;
  }
  __host__ __device__ static uint32_t done_parity(int B) { return (uint32_t)(B / kGjSlots) & 1u; }
};
}  // namespace isr
