// All-reduce of a few hundred doubles over NVLink peer memory, executed by ONE thread block inside a kernel that has other
// work to do (the last block of the chi2 test's range / cells kernels) -- the exchange step of the multi-GPU chi2 test fused
// into the kernels that produce and consume its operands, instead of two extra NCCL launches per test.
//
// Every rank owns a mailbox in its HBM that all peers have mapped (cudaIpc between processes, peer access inside one):
//     data  [2 parities][nranks][kPeerSlotDoubles]     slot [p][r] is written by rank r only
//     flags [2 parities][nranks]                       epoch of the last exchange rank r has delivered into parity p
// Exchange number e (the same on every rank; parity e & 1): each rank stores its values into slot [p][rank] of EVERY
// mailbox (its own included), fences system-wide, then stores e into flag [p][rank] of every mailbox; it then waits until
// all flags [p][*] of its own mailbox show e and reduces the nranks slots in RANK ORDER -- every rank computes the same
// bits, whatever the arrival order (NCCL gives no such guarantee across algorithms).  Two parities suffice: a rank cannot
// start exchange e + 2 before every peer has delivered e + 1, i.e. has finished reading e.
// A rank that never shows up must not hang the GPU: the wait gives up after about two seconds and raises *status.
#pragma once
#include <cstdint>

namespace isr {

constexpr int kPeerMaxRanks = 16;
constexpr int kPeerSlotDoubles = 12288;     // 96 kB per slot: TH1F cells (<= 8192) + 4 N + 2 fits
constexpr size_t kPeerDataDoubles = (size_t)2 * kPeerMaxRanks * kPeerSlotDoubles;
constexpr size_t kPeerBoxBytes = kPeerDataDoubles * sizeof(double) + (size_t)2 * kPeerMaxRanks * sizeof(unsigned long long);

struct PeerComm {               // passed to kernels by value
  double* box[kPeerMaxRanks];   // every rank's mailbox as mapped into THIS process / device
  int rank, nranks;
  int* status;                  // device word of this rank: set to 1 on a time-out
};

__device__ __forceinline__ unsigned long long* peer_flags(double* box) {
  return reinterpret_cast<unsigned long long*>(box + kPeerDataDoubles);
}

// op: 0 sum, 1 min.  vals / out: `len` doubles in this rank's memory (may alias).  Called by ALL threads of one block.
__device__ inline void peer_allreduce_block(const PeerComm& pc, unsigned long long epoch, const double* vals, double* out, int len, int op) {
  const int p = (int)(epoch & 1ull);
  const size_t slot = ((size_t)p * kPeerMaxRanks + pc.rank) * kPeerSlotDoubles;
  for (int r = 0; r < pc.nranks; ++r) {
    double* dst = pc.box[r] + slot;
    for (int i = threadIdx.x; i < len; i += blockDim.x) dst[i] = vals[i];
  }
  __threadfence_system();
  __syncthreads();
  if ((int)threadIdx.x < pc.nranks) {
    volatile unsigned long long* f = peer_flags(pc.box[threadIdx.x]) + p * kPeerMaxRanks + pc.rank;
    *f = epoch;                                           // deliver: my data for `epoch` is in rank threadIdx.x's mailbox
  }
  if ((int)threadIdx.x < pc.nranks) {
    volatile unsigned long long* f = peer_flags(pc.box[pc.rank]) + p * kPeerMaxRanks + threadIdx.x;
    const long long t0 = clock64();
    while (*f < epoch) {
      if (clock64() - t0 > 4000000000ll) { *pc.status = 1; break; }     // about two seconds: a peer is missing
    }
  }
  __threadfence_system();
  __syncthreads();
  const double* mine = pc.box[pc.rank] + (size_t)p * kPeerMaxRanks * kPeerSlotDoubles;
  for (int i = threadIdx.x; i < len; i += blockDim.x) {
    double acc = __ldcv(mine + i);                        // (volatile loads: the slots were written by other GPUs)
    for (int r = 1; r < pc.nranks; ++r) {
      const double v = __ldcv(mine + (size_t)r * kPeerSlotDoubles + i);
      acc = op == 0 ? acc + v : fmin(acc, v);
    }
    out[i] = acc;
  }
  __syncthreads();
}

}  // namespace isr
