// Host-ingress ceiling of a multi-GPU box: pure cudaMemcpyAsync from page-locked host memory to 1 / 2 / 4 / 8 GPUs at
// once, one host thread per GPU, with the page-locked buffers placed (a) wherever an unbound thread's cudaMallocHost puts
// them and (b) NUMA-local to the GPU (isr_host_alloc, which binds the allocating thread to the GPU's node first).
// the e2e curve of the chi2 test (1.68 GB of toys per GPU and step) fell to 0.42 at 8 GPUs; this probe
// records what the box can deliver, so that the e2e figure can be read against it.
//
//   nvcc -O2 -std=c++17 tools/h2d_probe.cu -Iinclude -Lisrsolver_b200 -lisr_b200 -Xlinker -rpath=$PWD/isrsolver_b200 -o tools/h2d_probe
//   tools/h2d_probe [bytes_per_gpu = 1677721600] [reps = 5]        -> one JSON object on stdout
#include <cuda_runtime.h>
#include <pthread.h>

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "isr_b200.h"

static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

struct Lane {
  int dev = 0, node = -1;
  isr_ctx* ctx = nullptr;
  void* host = nullptr;
  void* devbuf = nullptr;
  cudaStream_t st[2] = {nullptr, nullptr};
  double gbs = 0;
};

int main(int argc, char** argv) {
  const size_t bytes = argc > 1 ? strtoull(argv[1], nullptr, 10) : 1677721600ull;
  const int reps = argc > 2 ? atoi(argv[2]) : 5;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { fprintf(stderr, "no CUDA device\n"); return 1; }
  std::string out = "{\"bytes_per_gpu\": " + std::to_string(bytes) + ", \"reps\": " + std::to_string(reps) + ", \"gpus_visible\": " + std::to_string(ndev) + ", \"runs\": [";
  bool first = true;
  for (int local = 0; local <= 1; ++local) {
    if (local) unsetenv("ISR_NO_NUMA_BIND"); else setenv("ISR_NO_NUMA_BIND", "1", 1);
    for (int streams = 1; streams <= 2; ++streams) {
      for (int k = 1; k <= ndev; k *= 2) {
        std::vector<Lane> lanes(k);
        pthread_barrier_t bar;
        pthread_barrier_init(&bar, nullptr, k);
        std::vector<std::thread> th;
        double t_all0 = 0, t_all1 = 0;
        for (int r = 0; r < k; ++r)
          th.emplace_back([&, r] {
            Lane& L = lanes[r];
            L.dev = r;
            if (isr_ctx_create(r, &L.ctx) != ISR_OK) { fprintf(stderr, "ctx %d: %s\n", r, isr_last_error()); exit(1); }
            if (local) isr_ctx_bind_host_thread(L.ctx, &L.node);
            if (isr_host_alloc(L.ctx, bytes, &L.host) != ISR_OK) { fprintf(stderr, "alloc %d: %s\n", r, isr_last_error()); exit(1); }
            cudaSetDevice(r);
            cudaMalloc(&L.devbuf, bytes);
            for (int s = 0; s < streams; ++s) cudaStreamCreateWithFlags(&L.st[s], cudaStreamNonBlocking);
            auto copy_once = [&] {
              // `streams` concurrent copies of contiguous halves, each in 64 MiB pieces (2 MiB aligned)
              const size_t part = (bytes / streams) & ~((size_t)(2 << 20) - 1), piece = (size_t)64 << 20;
              for (int s = 0; s < streams; ++s) {
                const size_t b0 = (size_t)s * part, b1 = s == streams - 1 ? bytes : b0 + part;
                for (size_t o = b0; o < b1; o += piece)
                  cudaMemcpyAsync((char*)L.devbuf + o, (char*)L.host + o, std::min(piece, b1 - o), cudaMemcpyHostToDevice, L.st[s]);
              }
              for (int s = 0; s < streams; ++s) cudaStreamSynchronize(L.st[s]);
            };
            copy_once();   // warm-up
            pthread_barrier_wait(&bar);
            if (r == 0) t_all0 = now();
            const double t0 = now();
            for (int q = 0; q < reps; ++q) copy_once();
            const double t1 = now();
            L.gbs = (double)bytes * reps / (t1 - t0) * 1e-9;
            pthread_barrier_wait(&bar);
            if (r == 0) t_all1 = now();
            for (int s = 0; s < streams; ++s) cudaStreamDestroy(L.st[s]);
            cudaFree(L.devbuf);
            isr_host_free(L.host);
            isr_ctx_destroy(L.ctx);
          });
        for (auto& t : th) t.join();
        pthread_barrier_destroy(&bar);
        double agg = (double)bytes * reps * k / (t_all1 - t_all0) * 1e-9;
        char buf[512];
        std::string per;
        for (int r = 0; r < k; ++r) { snprintf(buf, sizeof(buf), "%s%.1f", r ? ", " : "", lanes[r].gbs); per += buf; }
        std::string nodes;
        for (int r = 0; r < k; ++r) { snprintf(buf, sizeof(buf), "%s%d", r ? ", " : "", lanes[r].node); nodes += buf; }
        snprintf(buf, sizeof(buf), "%s{\"numa_local\": %s, \"copy_streams\": %d, \"gpus\": %d, \"aggregate_gb_s\": %.1f, \"per_gpu_gb_s\": [%s], \"numa_node\": [%s]}",
                 first ? "" : ", ", local ? "true" : "false", streams, k, agg, per.c_str(), nodes.c_str());
        out += buf;
        first = false;
      }
    }
  }
  out += "]}";
  puts(out.c_str());
  return 0;
}
