// gpu_reconstruct.cpp -- TEST INFRASTRUCTURE ONLY.
//
// "Identical reconstructions" check (BASELINE.json configs[0] and [4]): runs the UNMODIFIED
// reference Reconstructor twice on the same collection -- once over the reference's own CPU
// black box (ShuntingYardParser::evaluate_pre, as benchmarks/benchmarks.cpp:22-38 does) and once
// over include/firefly_gpu/GpuBlackBox.hpp -- and compares the reconstructed rational functions
// term by term.  Built by `make -C oracle ref` into oracle/_ref/ against libfirefly_ref.so and
// firefly_b200/lib/libffgpu.so; run on the GPU box by tests/test_gpu_reconstruct.py.
//
//   gpu_reconstruct <file.m> <bunch> <threads> <var> [<var> ...]
//   exit 0 = identical, 1 = different, 2 = usage
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
}  // namespace firefly

using namespace firefly;

template <typename BB>
static std::vector<std::string> reconstruct_all(BB& bb, uint32_t n, uint32_t threads, uint32_t bunch,
                                                const std::vector<std::string>& vars, uint32_t* probes) {
  FFInt::set_new_prime(primes()[0]);
  bb.prime_changed_internal();
  Reconstructor<BB> rec(n, threads, bunch, bb, Reconstructor<BB>::SILENT);
  rec.enable_shift_scan();
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
  CpuBlackBox cpu(par);
  const std::vector<std::string> want = reconstruct_all(cpu, n, threads, bunch, vars, &pc);

  FFInt::set_new_prime(primes()[0]);
  GpuBlackBox gpu(file, vars);
  const std::vector<std::string> got = reconstruct_all(gpu, n, threads, bunch, vars, &pg);

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
