// h2d_peak.cu -- what the host link of this box can carry: plain cudaMemcpyAsync between PINNED host memory and the
// device, up, down and both at once (separate streams), on 1..N GPUs at the same time (one host thread per GPU).
// The end-to-end numbers of bench.py (ffgpu_eval with host buffers) are bounded by these, not by the kernels.
//   build/h2d_peak [--gpus N] [--mb 256] [--reps 10]      -> one JSON object on stdout
#include <cuda_runtime.h>

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

#define CK(x)                                                                        \
  do {                                                                               \
    cudaError_t e_ = (x);                                                            \
    if (e_ != cudaSuccess) {                                                         \
      std::fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_));                  \
      std::exit(1);                                                                  \
    }                                                                                \
  } while (0)

struct Dev {
  void *h_up = nullptr, *h_down = nullptr, *d_up = nullptr, *d_down = nullptr;
  cudaStream_t s_up = nullptr, s_down = nullptr;
};

// mode 0: up, 1: down, 2: both.  Returns seconds for `reps` copies of `bytes` per direction on every GPU at once.
static double run(std::vector<Dev>& devs, size_t bytes, int reps, int mode, bool wc) {
  (void)wc;
  std::vector<std::thread> th;
  std::vector<double> secs(devs.size(), 0.0);
  for (size_t g = 0; g < devs.size(); ++g)
    th.emplace_back([&, g] {
      CK(cudaSetDevice((int)g));
      Dev& d = devs[g];
      auto go = [&] {
        for (int r = 0; r < reps; ++r) {
          if (mode != 1) CK(cudaMemcpyAsync(d.d_up, d.h_up, bytes, cudaMemcpyHostToDevice, d.s_up));
          if (mode != 0) CK(cudaMemcpyAsync(d.h_down, d.d_down, bytes, cudaMemcpyDeviceToHost, d.s_down));
        }
        CK(cudaStreamSynchronize(d.s_up));
        CK(cudaStreamSynchronize(d.s_down));
      };
      go();      // warm-up
      const auto t0 = std::chrono::steady_clock::now();
      go();
      secs[g] = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    });
  for (auto& t : th) t.join();
  double mx = 0;
  for (double s : secs) mx = s > mx ? s : mx;
  return mx;
}

int main(int argc, char** argv) {
  int gpus = 1, reps = 10;
  size_t mb = 256;
  for (int i = 1; i < argc; ++i) {
    if (!std::strcmp(argv[i], "--gpus") && i + 1 < argc) gpus = std::atoi(argv[++i]);
    else if (!std::strcmp(argv[i], "--mb") && i + 1 < argc) mb = (size_t)std::atoi(argv[++i]);
    else if (!std::strcmp(argv[i], "--reps") && i + 1 < argc) reps = std::atoi(argv[++i]);
  }
  int count = 0;
  CK(cudaGetDeviceCount(&count));
  if (gpus > count) gpus = count;
  const size_t bytes = mb << 20;
  std::printf("{\"mb_per_copy\": %zu, \"reps\": %d, \"visible_gpus\": %d", mb, reps, count);
  for (int wc = 0; wc < 2; ++wc) {
    for (int n = 1; n <= gpus; n *= 2) {
      std::vector<Dev> devs(n);
      for (int g = 0; g < n; ++g) {
        CK(cudaSetDevice(g));
        // the upload source may be write-combined: not cached by the CPU, faster for the device to read
        CK(cudaHostAlloc(&devs[g].h_up, bytes, cudaHostAllocPortable | (wc ? cudaHostAllocWriteCombined : 0)));
        CK(cudaHostAlloc(&devs[g].h_down, bytes, cudaHostAllocPortable));
        std::memset(devs[g].h_up, 1, bytes);
        std::memset(devs[g].h_down, 1, bytes);
        CK(cudaMalloc(&devs[g].d_up, bytes));
        CK(cudaMalloc(&devs[g].d_down, bytes));
        CK(cudaStreamCreateWithFlags(&devs[g].s_up, cudaStreamNonBlocking));
        CK(cudaStreamCreateWithFlags(&devs[g].s_down, cudaStreamNonBlocking));
      }
      const double up = run(devs, bytes, reps, 0, wc), down = run(devs, bytes, reps, 1, wc), both = run(devs, bytes, reps, 2, wc);
      const double gb = (double)bytes * reps * n / 1e9;
      const char* tag = wc ? "_wc" : "";
      if (n == 1 && !wc)
        std::printf(", \"h2d_gbs\": %.2f, \"d2h_gbs\": %.2f, \"both_gbs_sum\": %.2f", gb / up, gb / down, 2 * gb / both);
      std::printf(", \"n%d%s\": {\"h2d_gbs\": %.2f, \"d2h_gbs\": %.2f, \"both_gbs_sum\": %.2f}", n, tag, gb / up, gb / down, 2 * gb / both);
      for (int g = 0; g < n; ++g) {
        CK(cudaSetDevice(g));
        cudaFreeHost(devs[g].h_up);
        cudaFreeHost(devs[g].h_down);
        cudaFree(devs[g].d_up);
        cudaFree(devs[g].d_down);
        cudaStreamDestroy(devs[g].s_up);
        cudaStreamDestroy(devs[g].s_down);
      }
    }
  }
  std::printf(", \"note\": \"GB/s summed over the GPUs running at the same time; both = up and down on separate streams, sum of the two directions; _wc = upload from write-combined pinned memory\"}\n");
  return 0;
}
