// Communicators (NCCL over NVLink 5 / NVSwitch) and NUMA placement for the multi-GPU entry points of the C ABI.
//
// The reference is single-threaded and single-process; north_star shards toys / lambdas over the GPUs of one
// box "with a single small NCCL gather of chi2 histograms and covariance partial sums".  The collectives themselves are
// two all-reduces of O(N) doubles per chi2 test (chi2_test.cu).
//
// NCCL is bound at RUN time (dlopen of libnccl.so.2 when the first communicator is made) rather than at link time: inside
// a Python process PyTorch has usually loaded its own bundled libnccl.so.2 already and the loader then hands back that
// very copy (two different NCCL builds in one process are not safe), while a plain C++ caller gets the system library;
// single-GPU users need no NCCL at all.
#include "comm.hpp"

#include <dlfcn.h>
#include <sched.h>
#include <unistd.h>

#include <cctype>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <mutex>
#include <sstream>

namespace isr {

// ---- the few NCCL entry points used, with the enum values of nccl.h (stable across NCCL 2.x) -------------------------
typedef struct ncclComm* ncclComm_t;
struct ncclUniqueId { char internal[128]; };
enum { kNcclInt32 = 2, kNcclUint32 = 3, kNcclInt64 = 4, kNcclFloat64 = 8 };
enum { kNcclSum = 0, kNcclMax = 2, kNcclMin = 3 };
struct NcclApi {
  int (*GetVersion)(int*);
  int (*GetUniqueId)(ncclUniqueId*);
  int (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int);
  int (*CommInitAll)(ncclComm_t*, int, const int*);
  int (*CommDestroy)(ncclComm_t);
  int (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t);
  int (*AllGather)(const void*, void*, size_t, int, ncclComm_t, cudaStream_t);
  const char* (*GetErrorString)(int);
  int (*GroupStart)();
  int (*GroupEnd)();
  void* handle = nullptr;
};
static NcclApi* nccl_api() {
  static NcclApi api;
  static std::once_flag once;
  static bool ok = false;
  std::call_once(once, [] {
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* nm : names) {
      api.handle = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
      if (api.handle) break;
    }
    if (!api.handle) return;
#define ISR_NCCL_SYM(field, name)                                   \
  *reinterpret_cast<void**>(&api.field) = dlsym(api.handle, name);  \
  if (!api.field) return;
    ISR_NCCL_SYM(GetVersion, "ncclGetVersion")
    ISR_NCCL_SYM(GetUniqueId, "ncclGetUniqueId")
    ISR_NCCL_SYM(CommInitRank, "ncclCommInitRank")
    ISR_NCCL_SYM(CommInitAll, "ncclCommInitAll")
    ISR_NCCL_SYM(CommDestroy, "ncclCommDestroy")
    ISR_NCCL_SYM(AllReduce, "ncclAllReduce")
    ISR_NCCL_SYM(AllGather, "ncclAllGather")
    ISR_NCCL_SYM(GetErrorString, "ncclGetErrorString")
    ISR_NCCL_SYM(GroupStart, "ncclGroupStart")
    ISR_NCCL_SYM(GroupEnd, "ncclGroupEnd")
#undef ISR_NCCL_SYM
    ok = true;
  });
  if (!ok) {
    const char* why = dlerror();
    set_error("NCCL is not available (dlopen libnccl.so.2: %s); multi-GPU entry points need it", why ? why : "symbol missing");
    return nullptr;
  }
  return &api;
}
#define ISR_NCCL(api, expr)                                                                              \
  do {                                                                                                   \
    int r__ = (expr);                                                                                    \
    if (r__ != 0) {                                                                                      \
      isr::set_error("NCCL error '%s' at %s:%d (%s)", (api)->GetErrorString(r__), __FILE__, __LINE__, #expr); \
      return ISR_ECUDA;                                                                                  \
    }                                                                                                    \
  } while (0)

// the mailbox exchange as a kernel of its own (callers that cannot fuse it into a producer kernel)
__global__ void __launch_bounds__(256) peer_allreduce_kernel(PeerComm pc, unsigned long long epoch, double* buf, int len, int op) {
  peer_allreduce_block(pc, epoch, buf, buf, len, op);
}

int comm_allreduce(isr_comm* cm, void* buf, int64_t count, int dtype, int op) {
  if (!cm || !buf || count < 0) { set_error("isr_comm_allreduce_dev: bad argument"); return ISR_EINVAL; }
  if (count == 0 || cm->nranks == 1) return ISR_OK;
  if (dtype == 0 && op <= 1 && comm_peer_fits(cm, count)) {      // small FP64 sum / min: NVLink mailboxes, rank-ordered
    peer_allreduce_kernel<<<1, 256, 0, cm->ctx->stream>>>(cm->peer, ++cm->epoch, static_cast<double*>(buf), (int)count, op); ISR_COUNT_LAUNCH();
    ISR_CUDA(cudaGetLastError());
    return ISR_OK;
  }
  static const int dt[4] = {kNcclFloat64, kNcclInt32, kNcclUint32, kNcclInt64};
  static const int ops[3] = {kNcclSum, kNcclMin, kNcclMax};
  if (dtype < 0 || dtype > 3 || op < 0 || op > 2) { set_error("isr_comm_allreduce_dev: unknown dtype / op"); return ISR_EINVAL; }
  NcclApi* api = nccl_api();
  if (!api) return ISR_ECUDA;
  ISR_NCCL(api, api->AllReduce(buf, buf, (size_t)count, dt[dtype], ops[op], (ncclComm_t)cm->nccl, cm->ctx->stream));
  return ISR_OK;
}

// ---- NUMA placement ---------------------------------------------------------------------------------------------------
// CPUs of the NUMA node a device hangs off, from sysfs (no libnuma in the image): the device's PCI address ->
// /sys/bus/pci/devices/<addr>/numa_node -> /sys/devices/system/node/node<k>/cpulist.
static bool parse_cpulist(const std::string& s, cpu_set_t* set) {
  CPU_ZERO(set);
  std::stringstream ss(s);
  std::string tok;
  bool any = false;
  while (std::getline(ss, tok, ',')) {
    int a = -1, b = -1;
    if (sscanf(tok.c_str(), "%d-%d", &a, &b) == 2) {
      for (int q = a; q <= b && q < CPU_SETSIZE; ++q) { CPU_SET(q, set); any = true; }
    } else if (sscanf(tok.c_str(), "%d", &a) == 1 && a >= 0 && a < CPU_SETSIZE) {
      CPU_SET(a, set); any = true;
    }
  }
  return any;
}
// Fall-back when sysfs does not know the node (virtual machines report numa_node = -1): NVML's ideal CPU affinity of
// the device -- what `nvidia-smi topo -m` prints -- bound at run time like NCCL.  Returns a pseudo node id (1000 + first
// CPU of the set) so that callers can tell devices with different affinities apart.
static int nvml_cpu_affinity(const char* bus, cpu_set_t* cpus) {
  static void* h = dlopen("libnvidia-ml.so.1", RTLD_NOW);
  if (!h) return -1;
  typedef int (*InitFn)();
  typedef int (*ByBusFn)(const char*, void**);
  typedef int (*AffFn)(void*, unsigned, unsigned long*);
  static InitFn init = (InitFn)dlsym(h, "nvmlInit_v2");
  static ByBusFn by_bus = (ByBusFn)dlsym(h, "nvmlDeviceGetHandleByPciBusId_v2");
  static AffFn aff = (AffFn)dlsym(h, "nvmlDeviceGetCpuAffinity");
  if (!init || !by_bus || !aff || init() != 0) return -1;
  void* dev = nullptr;
  if (by_bus(bus, &dev) != 0) return -1;
  unsigned long mask[CPU_SETSIZE / (8 * sizeof(unsigned long))] = {};
  const unsigned words = sizeof(mask) / sizeof(mask[0]);
  if (aff(dev, words, mask) != 0) return -1;
  CPU_ZERO(cpus);
  int first = -1;
  for (unsigned w = 0; w < words; ++w)
    for (unsigned b = 0; b < 8 * sizeof(unsigned long); ++b)
      if (mask[w] >> b & 1ul) { const int cpu = (int)(w * 8 * sizeof(unsigned long) + b); CPU_SET(cpu, cpus); if (first < 0) first = cpu; }
  return first < 0 ? -1 : 1000 + first;
}
int device_numa_node(int device, cpu_set_t* cpus) {
  char bus[32] = "";
  if (cudaDeviceGetPCIBusId(bus, sizeof(bus), device) != cudaSuccess) { cudaGetLastError(); return -1; }
  char busl[32];
  std::strcpy(busl, bus);
  for (char* q = busl; *q; ++q) *q = (char)tolower(*q);
  std::ifstream f(std::string("/sys/bus/pci/devices/") + busl + "/numa_node");
  int node = -1;
  std::string list;
  bool have = false;
  if ((f >> node) && node >= 0) {
    std::ifstream g("/sys/devices/system/node/node" + std::to_string(node) + "/cpulist");
    have = std::getline(g, list) && parse_cpulist(list, cpus);
  }
  if (!have) {
    node = nvml_cpu_affinity(bus, cpus);
    if (node < 0) return -1;
  }
  // keep only CPUs this process may run on (cgroup / taskset restrictions)
  cpu_set_t allowed;
  if (sched_getaffinity(0, sizeof(allowed), &allowed) == 0) {
    cpu_set_t both;
    CPU_AND(&both, cpus, &allowed);
    if (CPU_COUNT(&both) == 0) return -1;
    *cpus = both;
  }
  return node;
}
int bind_thread_to_device_node(int device) {
  if (const char* off = std::getenv("ISR_NO_NUMA_BIND")) { if (off[0] == '1') return -1; }
  cpu_set_t cpus;
  const int node = device_numa_node(device, &cpus);
  if (node < 0) return -1;
  if (sched_setaffinity(0, sizeof(cpus), &cpus) != 0) return -1;
  return node;
}
ScopedNodeBinding::ScopedNodeBinding(int device) {
  saved_ok = sched_getaffinity(0, sizeof(saved), &saved) == 0;
  node = bind_thread_to_device_node(device);
}
ScopedNodeBinding::~ScopedNodeBinding() {
  if (saved_ok && node >= 0) sched_setaffinity(0, sizeof(saved), &saved);
}

// ---- NVLink mailboxes ---------------------------------------------------------------------------------------------------
static bool peer_disabled() { const char* e = std::getenv("ISR_NO_P2P"); return e && e[0] == '1'; }

static int peer_alloc(isr_comm* cm) {
  ISR_CUDA(cudaSetDevice(cm->ctx->device));
  ISR_CUDA(cudaMalloc((void**)&cm->my_box, kPeerBoxBytes + sizeof(int)));
  ISR_CUDA(cudaMemset(cm->my_box, 0, kPeerBoxBytes + sizeof(int)));
  cm->peer.rank = cm->rank; cm->peer.nranks = cm->nranks;
  cm->peer.status = reinterpret_cast<int*>(reinterpret_cast<char*>(cm->my_box) + kPeerBoxBytes);
  ISR_CUDA(cudaMallocHost((void**)&cm->status_pin, sizeof(int)));
  *cm->status_pin = 0;
  return ISR_OK;
}
static void peer_release(isr_comm* cm) {
  if (!cm->my_box) return;
  cudaSetDevice(cm->ctx->device);
  if (cm->peer_ipc)
    for (int r = 0; r < cm->nranks; ++r)
      if (r != cm->rank && cm->peer.box[r]) cudaIpcCloseMemHandle(cm->peer.box[r]);
  cudaFree(cm->my_box);
  if (cm->status_pin) cudaFreeHost(cm->status_pin);
  cm->my_box = nullptr; cm->peer_ok = false;
}
// one process per GPU: mailboxes exchanged as CUDA IPC handles through ONE ncclAllGather; agreement on success through one
// ncclAllReduce(MIN) (a rank that cannot map a peer switches everybody back to NCCL)
static int peer_setup_ipc(isr_comm* cm, NcclApi* api) {
  if (cm->nranks > kPeerMaxRanks) return ISR_OK;
  int ok = peer_disabled() ? 0 : 1;
  if (ok && peer_alloc(cm) != ISR_OK) ok = 0;
  cudaStream_t st = cm->ctx->stream;
  cudaIpcMemHandle_t mine;
  std::memset(&mine, 0, sizeof(mine));
  if (ok && cudaIpcGetMemHandle(&mine, cm->my_box) != cudaSuccess) { cudaGetLastError(); ok = 0; }
  DevBuf<char> dsend, drecv;
  DevBuf<int> dflag;
  ISR_TRY(dsend.reserve(sizeof(mine))); ISR_TRY(drecv.reserve(sizeof(mine) * cm->nranks)); ISR_TRY(dflag.reserve(1));
  ISR_CUDA(cudaMemcpyAsync(dsend.p, &mine, sizeof(mine), cudaMemcpyHostToDevice, st));
  ISR_NCCL(api, api->AllGather(dsend.p, drecv.p, sizeof(mine), /*ncclInt8*/ 0, (ncclComm_t)cm->nccl, st));
  std::vector<cudaIpcMemHandle_t> all(cm->nranks);
  ISR_CUDA(cudaMemcpyAsync(all.data(), drecv.p, sizeof(mine) * cm->nranks, cudaMemcpyDeviceToHost, st));
  ISR_CUDA(cudaStreamSynchronize(st));
  if (ok) {
    for (int r = 0; r < cm->nranks && ok; ++r) {
      if (r == cm->rank) { cm->peer.box[r] = cm->my_box; continue; }
      void* ptr = nullptr;
      if (cudaIpcOpenMemHandle(&ptr, all[r], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); ok = 0; break; }
      cm->peer.box[r] = static_cast<double*>(ptr);
    }
    cm->peer_ipc = true;
  }
  ISR_CUDA(cudaMemcpyAsync(dflag.p, &ok, sizeof(int), cudaMemcpyHostToDevice, st));
  ISR_NCCL(api, api->AllReduce(dflag.p, dflag.p, 1, kNcclInt32, kNcclMin, (ncclComm_t)cm->nccl, st));     // also the barrier: all mailboxes zeroed and mapped
  int all_ok = 0;
  ISR_CUDA(cudaMemcpyAsync(&all_ok, dflag.p, sizeof(int), cudaMemcpyDeviceToHost, st));
  ISR_CUDA(cudaStreamSynchronize(st));
  cm->peer_ok = all_ok == 1;
  if (!cm->peer_ok) peer_release(cm);
  return ISR_OK;
}
// one process, several GPUs: peer access
static int peer_setup_local(int n, isr_comm** cms) {
  if (n > kPeerMaxRanks || peer_disabled()) return ISR_OK;
  bool ok = true;
  for (int r = 0; r < n && ok; ++r) ok = peer_alloc(cms[r]) == ISR_OK;
  for (int r = 0; r < n && ok; ++r) {
    cudaSetDevice(cms[r]->ctx->device);
    for (int q = 0; q < n && ok; ++q) {
      if (q == r) continue;
      int can = 0;
      cudaDeviceCanAccessPeer(&can, cms[r]->ctx->device, cms[q]->ctx->device);
      if (!can) { ok = false; break; }
      cudaError_t e = cudaDeviceEnablePeerAccess(cms[q]->ctx->device, 0);
      if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) ok = false;
      cudaGetLastError();
    }
  }
  for (int r = 0; r < n; ++r) {
    if (ok) {
      for (int q = 0; q < n; ++q) cms[r]->peer.box[q] = cms[q]->my_box;
      cms[r]->peer_ok = true;
    } else {
      peer_release(cms[r]);
    }
  }
  for (int r = 0; r < n; ++r) { cudaSetDevice(cms[r]->ctx->device); cudaDeviceSynchronize(); }
  return ISR_OK;
}

}  // namespace isr

using namespace isr;

extern "C" {

int isr_comm_unique_id(void* id_out) {
  if (!id_out) { set_error("isr_comm_unique_id: NULL"); return ISR_EINVAL; }
  static_assert(sizeof(ncclUniqueId) == ISR_COMM_ID_BYTES, "unique id size");
  NcclApi* api = nccl_api();
  if (!api) return ISR_ECUDA;
  ncclUniqueId id;
  ISR_NCCL(api, api->GetUniqueId(&id));
  std::memcpy(id_out, &id, sizeof(id));
  return ISR_OK;
}

int isr_comm_init_rank(isr_ctx* ctx, int nranks, int rank, const void* id, isr_comm** out) {
  if (!ctx || !out || nranks < 1 || rank < 0 || rank >= nranks || (nranks > 1 && !id)) { set_error("isr_comm_init_rank: bad argument"); return ISR_EINVAL; }
  *out = nullptr;
  isr_comm* cm = new (std::nothrow) isr_comm();
  if (!cm) { set_error("out of host memory"); return ISR_EINVAL; }
  cm->ctx = ctx; cm->rank = rank; cm->nranks = nranks;
  if (nranks > 1) {
    NcclApi* api = nccl_api();
    if (!api) { delete cm; return ISR_ECUDA; }
    ISR_CUDA(cudaSetDevice(ctx->device));
    ncclUniqueId uid;
    std::memcpy(&uid, id, sizeof(uid));
    ncclComm_t c = nullptr;
    int r = api->CommInitRank(&c, nranks, uid, rank);
    if (r != 0) { set_error("ncclCommInitRank failed: %s", api->GetErrorString(r)); delete cm; return ISR_ECUDA; }
    cm->nccl = c;
    const int rc = peer_setup_ipc(cm, api);
    if (rc != ISR_OK) { isr_comm_destroy(cm); return rc; }
  }
  *out = cm;
  return ISR_OK;
}

int isr_comm_init_all(int n, isr_ctx* const* ctxs, isr_comm** out) {
  if (n < 1 || !ctxs || !out) { set_error("isr_comm_init_all: bad argument"); return ISR_EINVAL; }
  for (int r = 0; r < n; ++r) {
    out[r] = nullptr;
    if (!ctxs[r]) { set_error("isr_comm_init_all: ctxs[%d] == NULL", r); return ISR_EINVAL; }
    for (int q = 0; q < r; ++q)
      if (ctxs[q]->device == ctxs[r]->device) { set_error("isr_comm_init_all: device %d appears twice", ctxs[r]->device); return ISR_EINVAL; }
  }
  std::vector<ncclComm_t> cs(n, nullptr);
  if (n > 1) {
    NcclApi* api = nccl_api();
    if (!api) return ISR_ECUDA;
    std::vector<int> devs(n);
    for (int r = 0; r < n; ++r) devs[r] = ctxs[r]->device;
    ISR_NCCL(api, api->CommInitAll(cs.data(), n, devs.data()));
  }
  for (int r = 0; r < n; ++r) {
    isr_comm* cm = new (std::nothrow) isr_comm();
    if (!cm) { set_error("out of host memory"); return ISR_EINVAL; }
    cm->ctx = ctxs[r]; cm->rank = r; cm->nranks = n; cm->nccl = cs[r];
    out[r] = cm;
  }
  if (n > 1) ISR_TRY(peer_setup_local(n, out));
  return ISR_OK;
}

void isr_comm_destroy(isr_comm* cm) {
  if (!cm) return;
  peer_release(cm);
  if (cm->nccl) {
    NcclApi* api = nccl_api();
    if (api) { cudaSetDevice(cm->ctx->device); api->CommDestroy((ncclComm_t)cm->nccl); }
  }
  delete cm;
}

int isr_comm_uses_peer_memory(const isr_comm* cm) { return (cm && cm->peer_ok) ? 1 : 0; }
int isr_comm_rank(const isr_comm* cm) { return cm ? cm->rank : -1; }
int isr_comm_size(const isr_comm* cm) { return cm ? cm->nranks : 0; }

int isr_comm_allreduce_dev(isr_comm* cm, void* buf_dev, int64_t count, int dtype, int op) {
  if (!cm) { set_error("isr_comm_allreduce_dev: comm == NULL"); return ISR_EINVAL; }
  ISR_CUDA(cudaSetDevice(cm->ctx->device));
  return comm_allreduce(cm, buf_dev, count, dtype, op);
}

int isr_ctx_bind_host_thread(isr_ctx* ctx, int* node_out) {
  if (!ctx) { set_error("ctx == NULL"); return ISR_EINVAL; }
  const int node = bind_thread_to_device_node(ctx->device);
  if (node_out) *node_out = node;
  return ISR_OK;
}

}  // extern "C"
