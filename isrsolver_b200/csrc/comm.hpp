// Communicator and NUMA helpers (comm.cu).
#pragma once
#include <sched.h>

#include "common.cuh"

#include "peer_device.cuh"

struct isr_comm {
  isr_ctx* ctx = nullptr;
  int rank = 0, nranks = 1;
  void* nccl = nullptr;   // ncclComm_t (nullptr for a single rank)
  // NVLink peer-memory mailboxes for the small exchanges (peer_device.cuh); peer.box[r] == nullptr: not available, NCCL is used
  isr::PeerComm peer{};
  bool peer_ok = false;
  bool peer_ipc = false;          // peers mapped with cudaIpcOpenMemHandle (one process per GPU) rather than peer access
  double* my_box = nullptr;       // this rank's mailbox (cudaMalloc)
  unsigned long long epoch = 0;   // exchanges issued so far (the same sequence on every rank)
  int* status_pin = nullptr;      // page-locked copy target of peer.status
};

namespace isr {
// in-place all-reduce on the communicator's context stream; dtype 0 f64 / 1 i32 / 2 u32 / 3 i64, op 0 sum / 1 min / 2 max
int comm_allreduce(isr_comm* cm, void* buf_dev, int64_t count, int dtype, int op);
// true if the mailbox path can carry `count` doubles
inline bool comm_peer_fits(const isr_comm* cm, int64_t count) { return cm && cm->peer_ok && count <= kPeerSlotDoubles; }
// NUMA node of a device (-1 unknown) and the CPUs of that node this process may use
int device_numa_node(int device, cpu_set_t* cpus);
int bind_thread_to_device_node(int device);
// binds the calling thread to the device's node for the lifetime of the object
struct ScopedNodeBinding {
  explicit ScopedNodeBinding(int device);
  ~ScopedNodeBinding();
  cpu_set_t saved;
  bool saved_ok = false;
  int node = -1;
};
}  // namespace isr
