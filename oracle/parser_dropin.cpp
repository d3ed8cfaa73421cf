// parser_dropin.cpp -- TEST INFRASTRUCTURE ONLY.
//
// firefly::GpuShuntingYardParser (include/firefly_gpu/GpuShuntingYardParser.hpp) against the UNMODIFIED reference
// ShuntingYardParser, method by method, on the same collection:
//
//   parser_dropin host <file.m> <var>...   host logic only (runs without a GPU): sizes, RPN strings, add_otf positions,
//                                          trim maps, parse_function on an empty parser
//   parser_dropin gpu  <file.m> <var>...   the above plus the values of evaluate_pre<FFInt>, <FFIntVec<4>> and
//                                          <FFIntVec<128>> over two primes, on copies, after add_otf and trim, and the
//                                          reference Reconstructor run over the black box of test/test_bunch.cpp:29-58
//                                          with ONE line changed (the type of its parser member)
//
// Built by `make -C oracle ref` into oracle/_ref/ against libfirefly_ref.so and firefly_b200/lib/libffgpu.so.
// exit 0 = identical, 1 = different, 2 = usage
#include <cstdio>
#include <cstdlib>
#include <algorithm>
#include <cstring>
#include <stdexcept>
#include <string>
#include <unordered_set>
#include <vector>

#include "firefly/DenseSolver.hpp"
#include "firefly/Reconstructor.hpp"
#include "firefly/ShuntingYardParser.hpp"
#include "firefly_gpu/GpuShuntingYardParser.hpp"

using namespace firefly;

static int bad = 0;
#define CHECK(cond, ...)                  \
  do {                                    \
    if (!(cond)) {                        \
      ++bad;                              \
      std::printf("MISMATCH %s:%d: ", __FILE__, __LINE__); \
      std::printf(__VA_ARGS__);           \
      std::printf("\n");                  \
    }                                     \
  } while (0)

namespace firefly {
// the black box of the reference's test/test_bunch.cpp:29-58, templated on the parser type only
template <typename Parser>
class BlackBoxUser : public BlackBoxBase<BlackBoxUser<Parser>> {
 public:
  BlackBoxUser(const Parser& par_, int abort_at_ = 0) : par(par_), abort_at(abort_at_) {}
  template <typename FFIntTemp>
  std::vector<FFIntTemp> operator()(const std::vector<FFIntTemp>& values) {
    std::vector<FFIntTemp> result = par.evaluate_pre(values);
    result.emplace_back(result[0] / result[3]);
    mat_ff<FFIntTemp> mat = {{result[0], result[1]}, {result[2], result[3]}};
    std::vector<int> p{};
    calc_lu_decomposition(mat, p, 2);
    result.emplace_back(calc_determinant_lu(mat, p, 2));
    return result;
  }
  inline void prime_changed() {
    par.precompute_tokens();
    // test/test_save_write.cpp:51-60: give up when a later prime starts, the states are on disk by then
    if (abort_at && ++c == abort_at) throw std::runtime_error("Abort for save test.");
  }

 private:
  Parser par;
  int abort_at = 0, c = 0;
};
}  // namespace firefly

static uint64_t rnd(uint64_t& s) {
  uint64_t z = (s += 0x9E3779B97F4A7C15ULL);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
  return z ^ (z >> 31);
}

template <typename A, typename B>
static void same_rpn(const A& a, const B& b, const char* what) {
  const auto ra = a.get_rp_functions();
  const auto rb = b.get_rp_functions();
  CHECK(ra.size() == rb.size(), "%s: %zu vs %zu functions", what, ra.size(), rb.size());
  for (size_t i = 0; i < ra.size() && i < rb.size(); ++i) CHECK(ra[i] == rb[i], "%s: RPN of function %zu differs", what, i);
  CHECK(a.get_size() == b.get_size() && a.empty() == b.empty(), "%s: get_size / empty", what);
}

template <typename T, typename A, typename B>
static void same_values(const A& cpu, const B& gpu, size_t n_vars, uint64_t seed, const char* what) {
  constexpr size_t lanes = sizeof(T) / sizeof(uint64_t);
  std::vector<T> values(n_vars);
  uint64_t* raw = reinterpret_cast<uint64_t*>(values.data());
  for (size_t i = 0; i < n_vars * lanes; ++i) raw[i] = 1 + rnd(seed) % (FFInt::p - 1);
  raw[0] = 0;                                           // a zero coordinate in lane 0
  const std::vector<T> want = cpu.evaluate_pre(values);
  const std::vector<T> got = gpu.evaluate_pre(values);
  CHECK(want.size() == got.size(), "%s: %zu vs %zu values", what, want.size(), got.size());
  if (want.size() == got.size())
    CHECK(std::memcmp(want.data(), got.data(), want.size() * sizeof(T)) == 0, "%s: values differ (%zu lanes)", what, lanes);
}

template <typename A, typename B>
static void same_values_all(const A& cpu, const B& gpu, size_t n_vars, const char* what) {
  same_values<FFInt>(cpu, gpu, n_vars, 11, what);
  same_values<FFIntVec<4>>(cpu, gpu, n_vars, 12, what);
  same_values<FFIntVec<128>>(cpu, gpu, n_vars, 13, what);
}

template <typename BB>
static std::vector<std::string> reconstruct_all(BB& bb, uint32_t n, const std::vector<std::string>& vars) {
  FFInt::set_new_prime(primes()[0]);
  bb.prime_changed_internal();
  Reconstructor<BB> rec(n, 4, 4, bb, Reconstructor<BB>::SILENT);       // the reference test's bunch size
  rec.enable_shift_scan();
  rec.reconstruct();
  std::vector<std::string> out;
  for (const auto& rf : rec.get_result()) out.push_back(rf.to_string(vars));
  RatReconst::reset();
  return out;
}

// RationalFunction::to_string lists the monomials in hash-map order, which differs between a resumed and an
// uninterrupted run: sort the '+'-separated terms of every top-level parenthesis.
static std::string canon(const std::string& f) {
  std::string out;
  std::vector<std::string> terms;
  std::string cur;
  int depth = 0;
  for (char ch : f) {
    if (ch == '(' && depth++ == 0) { out += ch; continue; }
    if (ch == ')' && --depth == 0) {
      terms.push_back(cur);
      cur.clear();
      std::sort(terms.begin(), terms.end());
      for (size_t i = 0; i < terms.size(); ++i) out += (i ? "+" : "") + terms[i];
      terms.clear();
      out += ch;
      continue;
    }
    if (depth == 0) out += ch;
    else if (depth == 1 && ch == '+') { terms.push_back(cur); cur.clear(); }
    else cur += ch;
  }
  return out;
}

// The reference's save / resume flow (test/test_save_write.cpp:139-157): a run over the CPU parser writes its states
// (ff_save/ in the working directory) and aborts when the second prime starts; a second Reconstructor resumes from
// them over `Resumed` -- which first has to reproduce ff_save/validation.gz, the reference's own known-answer check
// of the black box (Reconstructor.hpp:626-657) -- and must finish with the functions of the uninterrupted run.
template <typename Resumed>
static void save_and_resume(const std::string& file, const std::vector<std::string>& vars,
                            const std::vector<std::string>& want) {
  const uint32_t n = (uint32_t)vars.size();
  bool aborted = false;
  FFInt::set_new_prime(primes()[0]);
  try {
    ShuntingYardParser p(file, vars);
    BlackBoxUser<ShuntingYardParser> b(p, 2);
    Reconstructor<BlackBoxUser<ShuntingYardParser>> r(n, 4, 4, b, Reconstructor<BlackBoxUser<ShuntingYardParser>>::SILENT);
    r.enable_shift_scan();
    r.set_tags();
    r.reconstruct();
  } catch (std::exception& e) {
    aborted = true;
  }
  RatReconst::reset();
  CHECK(aborted, "save/resume: the first run finished in one prime, nothing was resumed");
  if (!aborted) return;
  Resumed p2(file, vars);
  BlackBoxUser<Resumed> b2(p2);
  Reconstructor<BlackBoxUser<Resumed>> r2(n, 4, 4, b2, Reconstructor<BlackBoxUser<Resumed>>::SILENT);
  r2.set_tags();
  r2.resume_from_saved_state();
  r2.reconstruct();
  std::vector<std::string> got;
  for (const auto& rf : r2.get_result()) got.push_back(rf.to_string(vars));
  RatReconst::reset();
  CHECK(want.size() == got.size(), "save/resume: %zu vs %zu functions", want.size(), got.size());
  for (size_t i = 0; i < want.size() && i < got.size(); ++i)
    CHECK(canon(want[i]) == canon(got[i]), "save/resume: function %zu differs:\n  %.300s\n  %.300s", i,
          canon(want[i]).c_str(), canon(got[i]).c_str());
  std::printf("resumed from the saved states: %zu functions\n", got.size());
}

int main(int argc, char** argv) {
  if (argc < 4 || (std::strcmp(argv[1], "host") && std::strcmp(argv[1], "gpu") && std::strcmp(argv[1], "resume-cpu"))) {
    std::fprintf(stderr, "usage: parser_dropin host|gpu|resume-cpu <file.m> <var>...\n");
    return 2;
  }
  if (!std::strcmp(argv[1], "resume-cpu")) {
    // the harness of the save / resume check with the reference parser on both sides (runs without a GPU)
    std::vector<std::string> vars;
    for (int i = 3; i < argc; ++i) vars.emplace_back(argv[i]);
    FFInt::set_new_prime(primes()[0]);
    ShuntingYardParser pc(argv[2], vars);
    BlackBoxUser<ShuntingYardParser> bc(pc);
    const std::vector<std::string> want = reconstruct_all(bc, (uint32_t)vars.size(), vars);
    save_and_resume<ShuntingYardParser>(argv[2], vars, want);
    std::printf("resume-cpu checks done, %d differ, gpu kernel launches 0\n", bad);
    return bad ? 1 : 0;
  }
  const bool gpu = !std::strcmp(argv[1], "gpu");
  const std::string file = argv[2];
  std::vector<std::string> vars;
  for (int i = 3; i < argc; ++i) vars.emplace_back(argv[i]);
  const size_t n = vars.size();

  FFInt::set_new_prime(primes()[0]);
  ShuntingYardParser cpu(file, vars, false, true);
  GpuShuntingYardParser dev(file, vars, false, true);
  same_rpn(cpu, dev, "constructed from a file");
  if (gpu) same_values_all(cpu, dev, n, "constructed from a file");

  // a second prime through precompute_tokens(), and back
  FFInt::set_new_prime(primes()[3]);
  cpu.precompute_tokens();
  dev.precompute_tokens();
  if (gpu) same_values_all(cpu, dev, n, "after a prime change");

  // copies are independent parsers (black boxes hold theirs by value)
  {
    ShuntingYardParser cpu2(cpu);
    GpuShuntingYardParser dev2(dev), dev3;
    dev3 = dev2;
    same_rpn(cpu2, dev3, "copy");
    if (gpu) same_values_all(cpu2, dev3, n, "copy");
  }

  // functions added on the fly
  const std::vector<std::string> extra = {vars[0] + "^2-3*" + vars[n - 1], "(" + vars[0] + "+" + vars[n - 1] + ")/(" + vars[0] + "-7)",
                                          "12345678901234567890123*" + vars[0]};
  for (const std::string& e : extra) {
    std::string a = e, b = e;
    const size_t pc = cpu.add_otf(a), pd = dev.add_otf(b);
    CHECK(pc == pd, "add_otf position %zu vs %zu", pc, pd);
  }
  cpu.precompute_tokens(true);
  dev.precompute_tokens(true);
  same_rpn(cpu, dev, "after add_otf");
  if (gpu) same_values_all(cpu, dev, n, "after add_otf");

  // trim: every second function and the last one
  std::unordered_set<size_t> keep;
  for (size_t i = 0; i < cpu.get_size(); i += 2) keep.insert(i);
  keep.insert(cpu.get_size() - 1);
  keep.insert(cpu.get_size() + 5);                      // does not exist
  const auto mc = cpu.trim(keep);
  const auto md = dev.trim(keep);
  CHECK(mc == md, "trim maps differ (%zu vs %zu entries)", mc.size(), md.size());
  same_rpn(cpu, dev, "after trim");
  if (gpu) same_values_all(cpu, dev, n, "after trim");

  // functions that arrive as RPN
  {
    const std::vector<std::string> rpn = {vars[0], vars[n - 1], "*", "3", "+"};
    const size_t pc = cpu.add_otf_precompute(rpn), pd = dev.add_otf_precompute(rpn);
    CHECK(pc == pd, "add_otf_precompute position %zu vs %zu", pc, pd);
    if (gpu) same_values_all(cpu, dev, n, "after add_otf_precompute");
  }

  // the default-constructed parser filled by parse_function (Reconstructor's factor parsers, Reconstructor.hpp:1873-1874)
  {
    ShuntingYardParser c2;
    GpuShuntingYardParser d2;
    for (const std::string& e : extra) {
      std::string a = e, b = e;
      c2.parse_function(a, vars);
      d2.parse_function(b, vars);
    }
    same_rpn(c2, d2, "parse_function");
    c2.precompute_tokens();
    d2.precompute_tokens();
    CHECK(d2.get_size() == extra.size(), "parse_function: size");
    if (gpu) same_values_all(c2, d2, n, "parse_function");
  }

  // check_is_equal: the duplicated collection returns every value at its position
  {
    std::vector<std::string> funs = extra;
    funs.push_back(extra[0]);
    funs.push_back(vars[n - 1] + "*(-3)+" + vars[0] + "^2");            // equal to extra[0] as a function
    ShuntingYardParser c3(funs, vars, true, true);
    GpuShuntingYardParser d3(funs, vars, true, true);
    CHECK(c3.get_size() == 3 && d3.get_size() == 3, "check_is_equal: %zu / %zu distinct functions", c3.get_size(), d3.get_size());
    CHECK(d3.get_n_out() == 5, "check_is_equal: outputs");
    if (gpu) same_values_all(c3, d3, n, "check_is_equal");
  }

  if (gpu && n == 4) {
    // the reference's own test setting (test/test_bunch.cpp:60-77): s_y_4_v.m, 4 threads, bunch size 4
    FFInt::set_new_prime(primes()[0]);
    ShuntingYardParser pc(file, vars);
    BlackBoxUser<ShuntingYardParser> bc(pc);
    const std::vector<std::string> want = reconstruct_all(bc, (uint32_t)n, vars);
    FFInt::set_new_prime(primes()[0]);
    GpuShuntingYardParser pd(file, vars);
    BlackBoxUser<GpuShuntingYardParser> bd(pd);
    const std::vector<std::string> got = reconstruct_all(bd, (uint32_t)n, vars);
    CHECK(want.size() == got.size(), "reconstruction: %zu vs %zu functions", want.size(), got.size());
    for (size_t i = 0; i < want.size() && i < got.size(); ++i) CHECK(want[i] == got[i], "reconstructed function %zu differs", i);
    std::printf("%zu functions reconstructed over the GPU parser\n", got.size());
    save_and_resume<GpuShuntingYardParser>(file, vars, want);
  }

  std::printf("%s checks done, %d differ, gpu kernel launches %llu\n", argv[1], bad, (unsigned long long)ffgpu_launch_count());
  return bad ? 1 : 0;
}
