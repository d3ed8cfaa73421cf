// gpu_reconstruct.cpp -- TEST INFRASTRUCTURE ONLY.
//
// "Identical reconstructions" check (BASELINE.json configs[0] and [4]): runs the UNMODIFIED
// reference Reconstructor twice on the same collection -- once over the reference's own CPU
// black box (ShuntingYardParser::evaluate_pre, as benchmarks/benchmarks.cpp:22-38 does) and once
// over include/firefly_gpu/GpuBlackBox.hpp -- and compares the reconstructed rational functions
// term by term.  Built by `make -C oracle ref` into oracle/_ref/ against libfirefly_ref.so and
// firefly_b200/lib/libffgpu.so; run on the GPU box by tests/test_gpu_reconstruct.py.
//
// It also reports, for both runs, the quantity the reference itself prints as "Average time of the black-box probe"
// (Reconstructor.hpp:3381-3411, 3438-3472: wall time around every bb.eval call, summed, divided by the number of
// probes), measured the same way by a wrapper around each black box.
//
//   gpu_reconstruct <file.m> <bunch> <threads> <var> [<var> ...]
//   exit 0 = identical, 1 = different, 2 = usage
//   FFGPU_RECONSTRUCT_MODE = shift (default: enable_shift_scan, test/test_normal.cpp:89, test_bunch.cpp:69) |
//                            safe (set_safe_interpolation, test/test_safe.cpp:80, test_bunch.cpp:84) |
//                            none (neither, test/test_normal.cpp:81)
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

#include "firefly/Reconstructor.hpp"
#include "firefly/ShuntingYardParser.hpp"
#include "firefly_gpu/GpuBlackBox.hpp"

namespace firefly {
class CpuBlackBox : public BlackBoxBase<CpuBlackBox> {
 public:
  explicit CpuBlackBox(const ShuntingYardParser& par_) : par(par_) {}
  template <typename FFIntTemp>
  std::vector<FFIntTemp> operator()(const std::vector<FFIntTemp>& values) { return par.evaluate_pre(values); }
  inline void prime_changed() { par.precompute_tokens(); }
 private:
  ShuntingYardParser par;
};
// wall time around every call of the wrapped black box and the probes it evaluated (lanes of the bunch type)
template <typename Inner>
class Timed : public BlackBoxBase<Timed<Inner>> {
 public:
  explicit Timed(Inner& inner_) : inner(inner_) {}
  template <typename FFIntTemp>
  std::vector<FFIntTemp> operator()(const std::vector<FFIntTemp>& values) {
    const auto t0 = std::chrono::high_resolution_clock::now();
    std::vector<FFIntTemp> r = inner(values);
    const auto t1 = std::chrono::high_resolution_clock::now();
    ns.fetch_add((uint64_t)std::chrono::duration_cast<std::chrono::nanoseconds>(t1 - t0).count());
    probes.fetch_add(sizeof(FFIntTemp) / sizeof(FFInt));
    calls.fetch_add(1);
    return r;
  }
  inline void prime_changed() { inner.prime_changed(); }
  double average_black_box_time() const { return probes.load() ? 1e-9 * (double)ns.load() / (double)probes.load() : 0.; }
  std::atomic<uint64_t> ns{0}, probes{0}, calls{0};
 private:
  Inner& inner;
};
}  // namespace firefly

using namespace firefly;

template <typename BB>
static std::vector<std::string> reconstruct_all(BB& bb, uint32_t n, uint32_t threads, uint32_t bunch,
                                                const std::vector<std::string>& vars, uint32_t* probes) {
  FFInt::set_new_prime(primes()[0]);
  bb.prime_changed_internal();
  Reconstructor<BB> rec(n, threads, bunch, bb, Reconstructor<BB>::SILENT);
  const char* mode_env = std::getenv("FFGPU_RECONSTRUCT_MODE");
  const std::string mode = mode_env && *mode_env ? mode_env : "shift";
  if (mode == "shift") rec.enable_shift_scan();
  else if (mode == "safe") rec.set_safe_interpolation();
  else if (mode != "none") {
    std::fprintf(stderr, "FFGPU_RECONSTRUCT_MODE must be shift, safe or none\n");
    std::exit(2);
  }
  rec.reconstruct();
  std::vector<std::string> out;
  for (const auto& rf : rec.get_result()) out.push_back(rf.to_string(vars));
  *probes = 0;
  RatReconst::reset();
  return out;
}

int main(int argc, char** argv) {
  if (argc < 5) {
    std::fprintf(stderr, "usage: gpu_reconstruct <file.m> <bunch> <threads> <var>...\n");
    return 2;
  }
  const std::string file = argv[1];
  const uint32_t bunch = (uint32_t)std::atoi(argv[2]), threads = (uint32_t)std::atoi(argv[3]);
  std::vector<std::string> vars;
  for (int i = 4; i < argc; ++i) vars.emplace_back(argv[i]);
  const uint32_t n = (uint32_t)vars.size();
  uint32_t pc = 0, pg = 0;

  FFInt::set_new_prime(primes()[0]);
  ShuntingYardParser par(file, vars);
  CpuBlackBox cpu_bb(par);
  Timed<CpuBlackBox> cpu(cpu_bb);
  const auto c0 = std::chrono::steady_clock::now();
  const std::vector<std::string> want = reconstruct_all(cpu, n, threads, bunch, vars, &pc);
  const double cpu_wall = std::chrono::duration<double>(std::chrono::steady_clock::now() - c0).count();

  FFInt::set_new_prime(primes()[0]);
  GpuBlackBox gpu_bb(file, vars);
  gpu_bb.set_bunch_size(bunch);
  Timed<GpuBlackBox> gpu(gpu_bb);
  const auto g0 = std::chrono::steady_clock::now();
  const std::vector<std::string> got = reconstruct_all(gpu, n, threads, bunch, vars, &pg);
  const double gpu_wall = std::chrono::duration<double>(std::chrono::steady_clock::now() - g0).count();
  uint64_t co[3];
  gpu_bb.coalesce_stats(co);
  std::printf("black box  cpu: %llu probes in %llu calls, average time of the black-box probe %.3e s, reconstruct() %.2f s\n",
              (unsigned long long)cpu.probes.load(), (unsigned long long)cpu.calls.load(), cpu.average_black_box_time(), cpu_wall);
  std::printf("black box  gpu: %llu probes in %llu calls, average time of the black-box probe %.3e s, reconstruct() %.2f s, "
              "%llu shared launches for %llu calls\n",
              (unsigned long long)gpu.probes.load(), (unsigned long long)gpu.calls.load(), gpu.average_black_box_time(), gpu_wall,
              (unsigned long long)co[0], (unsigned long long)co[1]);

  int bad = want.size() != got.size();
  for (size_t i = 0; i < want.size() && i < got.size(); ++i) {
    if (want[i] != got[i]) {
      ++bad;
      std::printf("function %zu differs:\n  cpu: %.200s\n  gpu: %.200s\n", i, want[i].c_str(), got[i].c_str());
    }
  }
  std::printf("%zu functions reconstructed, %d differ, gpu kernel launches %llu\n", want.size(), bad,
              (unsigned long long)ffgpu_launch_count());
  return bad ? 1 : 0;
}
