// mailbox_latency.cu -- what a resident kernel that polls mapped pinned host memory costs per request on this box
// (development measurement for the small-call path of ffgpu_eval).  One CTA spins on a sequence word in host memory,
// copies `words` 64-bit words from host-mapped input to device memory, writes `words_out` words back to host-mapped
// output and publishes the sequence number; the host measures the round trip.  Bounded: the kernel leaves after
// `max_spins` polls without work, the host gives up after 2 s.
#include <cuda_runtime.h>

#include <atomic>
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { std::fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_)); return 1; } } while (0)

struct Mailbox {
  volatile uint64_t req;     // host -> device: sequence number of the newest request
  uint64_t pad0[15];
  volatile uint64_t done;    // device -> host
  uint64_t pad1[15];
  volatile uint64_t quit;
};

__global__ void server(Mailbox* mb, const uint64_t* h_in, uint64_t* h_out, uint64_t* d_buf, int words, int words_out,
                       unsigned long long max_spins) {
  __shared__ uint64_t seen_s;
  uint64_t served = 0;
  unsigned long long spins = 0;
  for (;;) {
    if (threadIdx.x == 0) {
      uint64_t r;
      for (;;) {
        r = mb->req;
        if (r != served || mb->quit || ++spins > max_spins) break;
      }
      seen_s = (r != served) ? r : 0;
    }
    __syncthreads();
    const uint64_t r = seen_s;
    if (r == 0) return;
    // inputs come over PCIe in one coalesced sweep, results go back the same way
    uint64_t acc = 0;
    for (int i = threadIdx.x; i < words; i += blockDim.x) {
      const uint64_t v = h_in[i];
      d_buf[i] = v;
      acc += v;
    }
    for (int i = threadIdx.x; i < words_out; i += blockDim.x) h_out[i] = acc + r + i;
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
      mb->done = r;
      served = r;
      spins = 0;
    }
    served = r;
    __syncthreads();
  }
}

__global__ void tiny(uint64_t* out, const uint64_t* in, int words) {
  uint64_t acc = 0;
  for (int i = threadIdx.x; i < words; i += blockDim.x) acc += in[i];
  out[threadIdx.x] = acc;
}

int main(int argc, char** argv) {
  int reps = 2000;
  if (argc > 1) reps = std::atoi(argv[1]);
  Mailbox* mb;
  uint64_t *h_in, *h_out, *d_buf;
  CK(cudaHostAlloc(&mb, sizeof(Mailbox), cudaHostAllocMapped));
  std::memset((void*)mb, 0, sizeof(Mailbox));
  const int max_words = 128 * 20;
  CK(cudaHostAlloc(&h_in, max_words * 8, cudaHostAllocMapped));
  CK(cudaHostAlloc(&h_out, max_words * 8, cudaHostAllocMapped));
  CK(cudaMalloc(&d_buf, max_words * 8));
  for (int i = 0; i < max_words; ++i) h_in[i] = i;
  cudaStream_t s;
  CK(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
  std::printf("{");
  const int shapes[3][2] = {{5, 1}, {128 * 5, 128}, {128 * 20, 128}};
  uint64_t seq = 0;
  for (int sh = 0; sh < 3; ++sh) {
    const int words = shapes[sh][0], words_out = shapes[sh][1];
    mb->quit = 0;
    server<<<1, 128, 0, s>>>(mb, h_in, h_out, d_buf, words, words_out, 5000000ULL);
    CK(cudaGetLastError());
    double best = 1e30, sum = 0;
    int ok = 1;
    for (int r = 0; r < reps && ok; ++r) {
      const auto t0 = std::chrono::steady_clock::now();
      ++seq;
      std::atomic_thread_fence(std::memory_order_seq_cst);
      mb->req = seq;
      while (mb->done != seq) {
        if (std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() > 2.0) { ok = 0; break; }
      }
      const double us = std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t0).count();
      if (h_out[words_out - 1] == 0) ok = 0;
      if (r >= reps / 10) { sum += us; if (us < best) best = us; }
    }
    mb->quit = 1;
    CK(cudaStreamSynchronize(s));
    std::printf("%s\"mailbox_%dw_in_%dw_out\": {\"ok\": %d, \"best_us\": %.2f, \"mean_us\": %.2f}", sh ? ", " : "", words, words_out, ok,
                best, sum / (reps - reps / 10));
    // the same job as launch + copies + synchronize
    uint64_t *d_in2, *d_out2;
    CK(cudaMalloc(&d_in2, max_words * 8));
    CK(cudaMalloc(&d_out2, 128 * 8));
    double best2 = 1e30, sum2 = 0;
    for (int r = 0; r < reps; ++r) {
      const auto t0 = std::chrono::steady_clock::now();
      CK(cudaMemcpyAsync(d_in2, h_in, words * 8, cudaMemcpyHostToDevice, s));
      tiny<<<1, 128, 0, s>>>(d_out2, d_in2, words);
      CK(cudaMemcpyAsync(h_out, d_out2, words_out * 8, cudaMemcpyDeviceToHost, s));
      CK(cudaStreamSynchronize(s));
      const double us = std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t0).count();
      if (r >= reps / 10) { sum2 += us; if (us < best2) best2 = us; }
    }
    std::printf(", \"launch_copy_sync_%dw\": {\"best_us\": %.2f, \"mean_us\": %.2f}", words, best2, sum2 / (reps - reps / 10));
    // launch with results written by the kernel into mapped memory (the shipped small-call path)
    double best3 = 1e30, sum3 = 0;
    for (int r = 0; r < reps; ++r) {
      const auto t0 = std::chrono::steady_clock::now();
      CK(cudaMemcpyAsync(d_in2, h_in, words * 8, cudaMemcpyHostToDevice, s));
      tiny<<<1, 128, 0, s>>>(h_out, d_in2, words);
      CK(cudaStreamSynchronize(s));
      const double us = std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t0).count();
      if (r >= reps / 10) { sum3 += us; if (us < best3) best3 = us; }
    }
    std::printf(", \"launch_copy_in_zero_copy_out_%dw\": {\"best_us\": %.2f, \"mean_us\": %.2f}", words, best3, sum3 / (reps - reps / 10));
    cudaFree(d_in2);
    cudaFree(d_out2);
  }
  std::printf("}\n");
  return 0;
}
