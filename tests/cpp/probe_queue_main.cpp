// Runtime check of include/firefly_gpu/GpuProbeQueue.hpp -- TEST INFRASTRUCTURE (tests/test_gpu_headers.py builds and
// runs it on the GPU box).  Pushes single probes and a line of probes through the header's C++ face of ffgpu_queue_*
// and prints "index value_0 value_1 ..." per finished probe; the test compares the lines with the oracle.
// usage: probe_queue_main <prime> <n_single> <n_line>      functions and points are fixed below
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <map>
#include <vector>

#include "ffgpu.h"
#include "firefly_gpu/GpuProbeQueue.hpp"

static uint64_t next(uint64_t& s, uint64_t p) {          // splitmix64, reduced into [1, p-1]
  uint64_t z = (s += 0x9E3779B97F4A7C15ULL);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
  z ^= z >> 31;
  return 1 + z % (p - 1);
}

int main(int argc, char** argv) {
  if (argc < 4) return 2;
  const uint64_t p = std::strtoull(argv[1], nullptr, 10);
  const size_t n_single = std::strtoull(argv[2], nullptr, 10), n_line = std::strtoull(argv[3], nullptr, 10);
  const char* funs[] = {"(x+y)^3/(z-x*y)", "x^7*y-3*z", "x/(y-y)", "12345678901234567890123*z^2+x"};
  const char* vars[] = {"x", "y", "z"};
  ffgpu_ctx* ctx = ffgpu_create(nullptr, 0);
  if (!ctx) { std::fprintf(stderr, "%s\n", ffgpu_last_error()); return 1; }
  ffgpu_tape* tape = ffgpu_compile(ctx, funs, 4, vars, 3, 0);
  if (!tape || ffgpu_set_prime(tape, p)) { std::fprintf(stderr, "%s\n", ffgpu_last_error()); return 1; }
  {
    firefly::GpuProbeQueue q(tape, 0);
    uint64_t s = 42;
    std::map<uint64_t, std::vector<uint64_t>> got;
    for (size_t i = 0; i < n_single; ++i) {
      std::vector<uint64_t> pt = {next(s, p), next(s, p), next(s, p)};
      if (q.push(pt) != i) { std::fprintf(stderr, "unexpected index\n"); return 1; }
    }
    std::vector<uint64_t> a = {1, next(s, p), next(s, p)}, sh = {next(s, p), next(s, p), next(s, p)}, t(n_line);
    for (auto& x : t) x = next(s, p);
    if (n_line && q.push_line(a, sh, t) != n_single) { std::fprintf(stderr, "unexpected line index\n"); return 1; }
    while (got.size() < n_single + n_line) {
      auto res = q.pop(97, true);
      if (res.first.empty()) break;
      for (size_t j = 0; j < res.first.size(); ++j) {
        std::vector<uint64_t> v;
        for (size_t f = 0; f < res.second.size(); ++f) v.push_back(res.second[f][j]);
        got[res.first[j]] = v;
      }
    }
    if (q.pending() != 0) { std::fprintf(stderr, "probes still pending\n"); return 1; }
    for (const auto& kv : got) {
      std::printf("%llu", (unsigned long long)kv.first);
      for (uint64_t v : kv.second) std::printf(" %llu", (unsigned long long)v);
      std::printf("\n");
    }
  }
  ffgpu_tape_destroy(tape);
  ffgpu_destroy(ctx);
  return 0;
}
