// isr_group: the GPUs of one box driven from ONE process (one host thread, one context, one replicated problem per
// device, one NCCL communicator over all of them).  This is what a C++ caller of chi2TestModel / ratioTestModel
// (include/Chi2Test.hpp:18-42, include/RatioTest.hpp:14-15; drivers src/isrsolver-SLE-chi2-test.cpp:106-157) links against
// to spread a test over a box -- north_star keeps the host in C++ with "a single small NCCL gather".
#include "comm.hpp"
#include "problem.hpp"
#include "solve.hpp"

#include <sys/mman.h>

#include <condition_variable>
#include <cstring>
#include <functional>
#include <map>
#include <mutex>
#include <string>
#include <thread>

namespace isr {

// One worker per device: runs the tasks posted to it, bound to the CPUs of its GPU's NUMA node.
struct Worker {
  std::thread th;
  std::mutex mu;
  std::condition_variable cv;
  std::function<int()> task;
  bool has_task = false, done = false, quit = false;
  int rc = ISR_OK;
  std::string err;
  int device = 0;
  void loop() {
    cudaSetDevice(device);
    bind_thread_to_device_node(device);
    std::unique_lock<std::mutex> lk(mu);
    for (;;) {
      cv.wait(lk, [this] { return has_task || quit; });
      if (quit) return;
      std::function<int()> t = std::move(task);
      has_task = false;
      lk.unlock();
      const int r = t();
      std::string e = r == ISR_OK ? std::string() : std::string(get_error());
      lk.lock();
      rc = r; err = std::move(e); done = true;
      cv.notify_all();
    }
  }
  void post(std::function<int()> t) {
    std::lock_guard<std::mutex> lk(mu);
    task = std::move(t); has_task = true; done = false;
    cv.notify_all();
  }
  int wait() {
    std::unique_lock<std::mutex> lk(mu);
    cv.wait(lk, [this] { return done; });
    return rc;
  }
};

}  // namespace isr

struct isr_group {
  int n = 0;
  std::vector<isr_ctx*> ctx;
  std::vector<isr_comm*> comm;
  std::vector<isr::Worker*> workers;
  std::map<double*, size_t> host_bufs;   // isr_group_host_alloc: pointer -> bytes
  // run f(rank) on every worker; the first failure's code and message are returned to the calling thread
  int run_all(const std::function<int(int)>& f) {
    for (int r = 0; r < n; ++r) workers[r]->post([&f, r] { return f(r); });
    int rc = ISR_OK;
    for (int r = 0; r < n; ++r) {
      const int q = workers[r]->wait();
      if (q != ISR_OK && rc == ISR_OK) { rc = q; isr::set_error("[gpu %d] %s", ctx[r] ? ctx[r]->device : r, workers[r]->err.c_str()); }
    }
    return rc;
  }
};
struct isr_gproblem {
  isr_group* g = nullptr;
  std::vector<isr_problem*> p;
  int n = 0;
};

using namespace isr;

static void shard_bounds(int64_t total, int world, int rank, int64_t* b, int64_t* e) {
  const int64_t base = total / world, rem = total % world;
  *b = rank * base + std::min<int64_t>(rank, rem);
  *e = *b + base + (rank < rem ? 1 : 0);
}

extern "C" {

int isr_group_create(int ngpu, const int* devices, isr_group** out) {
  if (!out || ngpu < 1) { set_error("isr_group_create: bad argument"); return ISR_EINVAL; }
  *out = nullptr;
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0) { set_error("isr_group_create: no CUDA device (%s); libisr_b200 has no CPU fallback", cudaGetErrorString(e)); return ISR_ECUDA; }
  if (ngpu > count) { set_error("isr_group_create: %d GPUs requested, %d visible", ngpu, count); return ISR_EINVAL; }
  isr_group* g = new (std::nothrow) isr_group();
  if (!g) { set_error("out of host memory"); return ISR_EINVAL; }
  g->n = ngpu;
  g->ctx.assign(ngpu, nullptr);
  g->comm.assign(ngpu, nullptr);
  for (int r = 0; r < ngpu; ++r) {
    Worker* w = new Worker();
    w->device = devices ? devices[r] : r;
    w->th = std::thread([w] { w->loop(); });
    g->workers.push_back(w);
  }
  int rc = g->run_all([g](int r) { return isr_ctx_create(g->workers[r]->device, &g->ctx[r]); });
  if (rc == ISR_OK) rc = isr_comm_init_all(ngpu, g->ctx.data(), g->comm.data());
  if (rc != ISR_OK) { isr_group_destroy(g); return rc; }
  *out = g;
  return ISR_OK;
}

void isr_group_destroy(isr_group* g) {
  if (!g) return;
  for (auto& kv : g->host_bufs) { cudaHostUnregister(kv.first); munmap(kv.first, kv.second); }
  for (isr_comm* cm : g->comm) isr_comm_destroy(cm);
  for (int r = 0; r < g->n; ++r) {
    Worker* w = g->workers[r];
    isr_ctx* c = g->ctx[r];
    if (c) { w->post([c] { isr_ctx_destroy(c); return ISR_OK; }); w->wait(); }
    { std::lock_guard<std::mutex> lk(w->mu); w->quit = true; w->cv.notify_all(); }
    w->th.join();
    delete w;
  }
  delete g;
}

int isr_group_size(const isr_group* g) { return g ? g->n : 0; }
isr_ctx* isr_group_ctx(isr_group* g, int rank) { return (g && rank >= 0 && rank < g->n) ? g->ctx[rank] : nullptr; }

int isr_group_problem_create(isr_group* g, int n, const double* ecm, double threshold, int nranges, const int* ranges,
                             const isr_eps_desc* eps, isr_gproblem** out) {
  if (!g || !out) { set_error("isr_group_problem_create: NULL argument"); return ISR_EINVAL; }
  *out = nullptr;
  isr_gproblem* gp = new (std::nothrow) isr_gproblem();
  if (!gp) { set_error("out of host memory"); return ISR_EINVAL; }
  gp->g = g; gp->n = n;
  gp->p.assign(g->n, nullptr);
  const int rc = g->run_all([&](int r) { return isr_problem_create(g->ctx[r], n, ecm, threshold, nranges, ranges, eps, &gp->p[r]); });
  if (rc != ISR_OK) { isr_group_problem_destroy(gp); return rc; }
  *out = gp;
  return ISR_OK;
}

void isr_group_problem_destroy(isr_gproblem* gp) {
  if (!gp) return;
  gp->g->run_all([gp](int r) { isr_problem_destroy(gp->p[r]); return ISR_OK; });
  delete gp;
}

isr_problem* isr_group_problem_member(isr_gproblem* gp, int rank) {
  return (gp && rank >= 0 && rank < (int)gp->p.size()) ? gp->p[rank] : nullptr;
}

int isr_group_host_alloc(isr_group* g, int64_t n_toys, int n, double** out) {
  if (!g || !out || n_toys < 0 || n < 1) { set_error("isr_group_host_alloc: bad argument"); return ISR_EINVAL; }
  *out = nullptr;
  const size_t bytes = std::max<size_t>(sizeof(double) * (size_t)n_toys * n, 4096);
  void* mem = mmap(nullptr, bytes, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
  if (mem == MAP_FAILED) { set_error("isr_group_host_alloc: mmap of %zu bytes failed", bytes); return ISR_EINVAL; }
  double* base = static_cast<double*>(mem);
  // first touch of every GPU's slice from that GPU's (node-bound) host thread: the pages land on its NUMA node
  g->run_all([&](int r) {
    int64_t b, e;
    shard_bounds(n_toys, g->n, r, &b, &e);
    if (e > b) std::memset(base + (size_t)b * n, 0, sizeof(double) * (size_t)(e - b) * n);
    return ISR_OK;
  });
  cudaError_t err = cudaHostRegister(mem, bytes, cudaHostRegisterPortable);
  if (err != cudaSuccess) {
    set_error("isr_group_host_alloc: cudaHostRegister failed (%s)", cudaGetErrorString(err));
    munmap(mem, bytes);
    return ISR_ECUDA;
  }
  g->host_bufs[base] = bytes;
  *out = base;
  return ISR_OK;
}

void isr_group_host_free(isr_group* g, double* ptr) {
  if (!g || !ptr) return;
  auto it = g->host_bufs.find(ptr);
  if (it == g->host_bufs.end()) return;
  cudaHostUnregister(ptr);
  munmap(ptr, it->second);
  g->host_bufs.erase(it);
}

int isr_group_chi2_test(isr_gproblem* gp, const isr_chi2_test_args* args) {
  if (!gp || !args) { set_error("isr_group_chi2_test: NULL argument"); return ISR_EINVAL; }
  if (args->comm || args->V_dev) { set_error("isr_group_chi2_test: comm and V_dev must be NULL (the group owns the communicators)"); return ISR_EINVAL; }
  isr_group* g = gp->g;
  const int n = gp->n;
  return g->run_all([&](int r) {
    isr_chi2_test_args a = *args;
    int64_t b, e;
    shard_bounds(args->n_toys, g->n, r, &b, &e);
    a.n_toys = e - b;
    a.first_toy = args->first_toy + b;
    if (args->V) a.V = args->V + (size_t)b * n;
    if (args->chi2_out) a.chi2_out = args->chi2_out + b;
    a.comm = g->comm[r];
    TestWants w;         // what is computed is the same on every rank ...
    w.sum = args->sum_bcs_out != nullptr; w.rhist = args->ratio_hist_out != nullptr; w.range = args->chi2_lo_hi_out != nullptr;
    if (r != 0) {        // ... but the global results are identical everywhere: rank 0 writes them
      a.chi2_lo_hi_out = nullptr; a.chi2_counts_out = nullptr; a.sum_bcs_out = nullptr; a.ratio_stats_out = nullptr;
      a.ratio_hist_out = nullptr; a.total_toys_out = nullptr; a.refactored_out = nullptr;
    }
    if (cudaSetDevice(gp->p[r]->ctx->device) != cudaSuccess) { set_error("cudaSetDevice failed"); return ISR_ECUDA; }
    return chi2_test(gp->p[r], a, w);
  });
}

// Condition number versus beam-energy spread (src/isrsolver-condnum-test.cpp:88-140) with the K settings cut into
// contiguous slices, one per GPU: no exchange at all (every GPU assembles the unspread matrix itself).
int isr_group_cond_spread_scan(isr_gproblem* gp, int K, const double* sigma, int per_point, double* cond_out, double* sv_out) {
  if (!gp || !sigma || !cond_out || K < 0) { set_error("isr_group_cond_spread_scan: bad argument"); return ISR_EINVAL; }
  isr_group* g = gp->g;
  const int n = gp->n;
  return g->run_all([&](int r) {
    int64_t b, e;
    shard_bounds(K, g->n, r, &b, &e);
    if (e == b) return (int)ISR_OK;
    const size_t per = per_point ? (size_t)n : 1;
    return isr_cond_spread_scan(gp->p[r], (int)(e - b), sigma + (size_t)b * per, per_point, cond_out + b,
                                sv_out ? sv_out + (size_t)b * n : nullptr);
  });
}

// isr_tikhonov_chi2_scan over all GPUs of the group: args->n_toys is the TOTAL and args->V the whole host array; every
// rank takes a contiguous slice of the toys, the per-lambda summaries returned are global (written by rank 0).
int isr_group_tikhonov_chi2_scan(isr_gproblem* gp, const isr_tikhonov_scan_args* args) {
  if (!gp || !args) { set_error("isr_group_tikhonov_chi2_scan: NULL argument"); return ISR_EINVAL; }
  if (args->comm || args->V_dev) { set_error("isr_group_tikhonov_chi2_scan: comm and V_dev must be NULL (the group owns the communicators)"); return ISR_EINVAL; }
  isr_group* g = gp->g;
  const int n = gp->n;
  return g->run_all([&](int r) {
    isr_tikhonov_scan_args a = *args;
    int64_t b, e;
    shard_bounds(args->n_toys, g->n, r, &b, &e);
    a.n_toys = e - b;
    if (args->V) a.V = args->V + (size_t)b * n;
    a.comm = g->comm[r];
    if (r != 0) { a.stats_out = nullptr; a.counts_out = nullptr; a.total_toys_out = nullptr; }
    return isr_tikhonov_chi2_scan(gp->p[r], &a);
  });
}

}  // extern "C"
