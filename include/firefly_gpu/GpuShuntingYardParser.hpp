// GpuShuntingYardParser.hpp -- header-only C++ adaptor: FireFly's ShuntingYardParser with the evaluation on B200s.
//
// Same public interface as the reference class (source/include/firefly/ShuntingYardParser.hpp:34-138), so that a
// user black box changes ONE line -- the type of its parser member:
//
//     class BlackBoxUser : public BlackBoxBase<BlackBoxUser> {          // test/test_bunch.cpp:29-58
//       BlackBoxUser(const GpuShuntingYardParser& par_) : par(par_) {}
//       template<typename FFIntTemp> std::vector<FFIntTemp> operator()(const std::vector<FFIntTemp>& values)
//         { return par.evaluate_pre(values); }
//       void prime_changed() { par.precompute_tokens(); }
//       GpuShuntingYardParser par;                                      // was: ShuntingYardParser par;
//     };
//
// Method by method:
//   ctor(file | funs, vars, check_is_equal, keep_rpn)  .cpp:32-58     ffgpu_compile_file / ffgpu_compile + precompute
//   evaluate_pre<FFInt | FFIntVec<N>>(values)          .hpp:365-479   ffgpu_eval, zero copy (see GpuBlackBox.hpp)
//   evaluate<T>(values), evaluate_pre_2<T>(values)     .hpp:240-363, 481-591   the same values, the same call
//   precompute_tokens(force)                           .cpp:550-591   ffgpu_set_prime(FFInt::p)
//   parse_function, add_otf, add_otf_precompute        .cpp:315-377   ffgpu_add_function[_rpn]
//   trim                                               .cpp:720-776   ffgpu_trim
//   get_rp_functions, get_rp_function, get_rp_functions_ref, move_rpn, empty, get_size, reserve   .cpp:343-345, 526-548
//
// Where it is deliberately kinder than the reference: the RPN strings and get_size() stay available after
// precompute_tokens() whatever keep_rpn says (the reference empties `functions` unless keep_rpn, so get_size() drops
// to 0 and trim() removes everything, .cpp:571-583, 735-741); add_otf on a check_is_equal parser works after
// construction (the reference has cleared its check points by then, .cpp:557-559).
//
// When the tape is rebuilt: parse_function and add_otf only record the function -- precompute_tokens() builds, as in
// the reference (an evaluation before that is an error here; the reference leaves the new functions out).  trim and
// add_otf_precompute leave a parser that evaluates at once, as in the reference, so they rebuild themselves.
//
// Copies are independent (ffgpu_tape_clone) and share one process-wide GPU context, created at first use over the
// current CUDA device -- or the devices given to GpuShuntingYardParser::use_devices() before that.
//
// Without a GPU the parser can still be built and inspected (RPN, sizes, trim maps: host logic); evaluate_pre then
// fails loudly.  There is no CPU evaluation path.  Errors: ERROR_MSG + std::exit(EXIT_FAILURE) like the reference.
#pragma once
#include <algorithm>
#include <cstdint>
#include <cstdlib>
#include <functional>
#include <memory>
#include <sstream>
#include <string>
#include <atomic>
#include <thread>
#include <unordered_map>
#include <unordered_set>
#include <vector>

#include "firefly/FFInt.hpp"
#include "firefly/FFIntVec.hpp"
#include "firefly/Logger.hpp"

#include "../ffgpu.h"
#include "caller_slot.hpp"

namespace firefly {

class GpuShuntingYardParser {
 public:
  GpuShuntingYardParser() {}
  GpuShuntingYardParser(const std::string& file, const std::vector<std::string>& vars, bool check_is_equal_ = false,
                        bool keep_rpn_ = false)
      : vars_(vars), check_is_equal(check_is_equal_), keep_rpn(keep_rpn_) {
    std::vector<const char*> v = c_strs(vars_);
    tape_ = ffgpu_compile_file(ctx(), file.c_str(), v.data(), v.size(), flags());
    if (!tape_) die(ffgpu_last_error());
    precompute_tokens();
  }
  GpuShuntingYardParser(const std::vector<std::string>& funs, const std::vector<std::string>& vars,
                        bool check_is_equal_ = false, bool keep_rpn_ = false)
      : vars_(vars), check_is_equal(check_is_equal_), keep_rpn(keep_rpn_) {
    std::vector<const char*> f = c_strs(funs), v = c_strs(vars_);
    tape_ = ffgpu_compile(ctx(), f.data(), f.size(), v.data(), v.size(), flags());
    if (!tape_) die(ffgpu_last_error());
    precompute_tokens();
  }
  GpuShuntingYardParser(const GpuShuntingYardParser& o)
      : vars_(o.vars_), check_is_equal(o.check_is_equal), keep_rpn(o.keep_rpn), prime_internal(o.prime_internal) {
    if (o.tape_ && !(tape_ = ffgpu_tape_clone(o.tape_))) die(ffgpu_last_error());
  }
  GpuShuntingYardParser(GpuShuntingYardParser&& o) noexcept { swap(o); }
  GpuShuntingYardParser& operator=(GpuShuntingYardParser o) noexcept {
    swap(o);
    return *this;
  }
  ~GpuShuntingYardParser() {
    if (tape_) ffgpu_tape_destroy(tape_);
  }

  // GPUs of the shared context; call before the first parser is built.  Empty = the current device.
  static void use_devices(const std::vector<int>& devices) { device_list() = devices; }

  // --- building ------------------------------------------------------------------------------------------------
  void parse_function(std::string& fun, const std::vector<std::string>& vars, bool validate_fun = false) {
    if (!tape_) {                       // the variables are fixed by the first call (vars_map.empty(), .cpp:318)
      vars_ = vars;
      std::vector<const char*> v = c_strs(vars_);
      tape_ = ffgpu_compile(ctx(), nullptr, 0, v.data(), v.size(), flags());
      if (!tape_) die(ffgpu_last_error());
    }
    if (ffgpu_add_function(tape_, fun.c_str(), FFGPU_ADD_PLAIN | (validate_fun ? FFGPU_ADD_VALIDATE : 0u)) < 0)
      die(ffgpu_last_error());
    rpn_valid_ = false;
  }
  size_t add_otf(std::string& fun, const bool no_duplicates = false) {
    need_tape("add_otf");
    const long pos = ffgpu_add_function(tape_, fun.c_str(), no_duplicates ? FFGPU_ADD_NO_DUPLICATES : 0u);
    if (pos < 0) die(ffgpu_last_error());
    rpn_valid_ = false;
    return static_cast<size_t>(pos);
  }
  size_t add_otf_precompute(const std::vector<std::string>& rpn_fun) {
    need_tape("add_otf_precompute");
    std::vector<const char*> t = c_strs(rpn_fun);
    const long pos = ffgpu_add_function_rpn(tape_, t.data(), t.size());
    if (pos < 0) die(ffgpu_last_error());
    rpn_valid_ = false;
    if (prime_internal) precompute_tokens(true);
    return static_cast<size_t>(pos);
  }
  void reserve(size_t) {}
  std::unordered_map<size_t, size_t> trim(const std::unordered_set<size_t>& elements_to_keep) {
    std::unordered_map<size_t, size_t> new_positions;
    if (!tape_) return new_positions;
    INFO_MSG("Trimming parser: " + std::to_string(get_size() - std::min(get_size(), elements_to_keep.size())) +
             " out of " + std::to_string(get_size()) + " functions will be removed");
    std::vector<uint64_t> keep(elements_to_keep.begin(), elements_to_keep.end()), np(elements_to_keep.size());
    if (ffgpu_trim(tape_, keep.data(), keep.size(), np.data())) die(ffgpu_last_error());
    for (size_t k = 0; k < keep.size(); ++k)
      if (np[k] != ~0ULL) new_positions.emplace(static_cast<size_t>(keep[k]), static_cast<size_t>(np[k]));
    rpn_valid_ = false;
    if (prime_internal) precompute_tokens(true);
    return new_positions;
  }

  // --- prime ---------------------------------------------------------------------------------------------------
  // Call after FFInt::set_new_prime, never concurrently with evaluate_pre (the reference's rule as well).
  void precompute_tokens(bool force = false) {
    if (!tape_) return;
    if (FFInt::p == prime_internal && !force && !ffgpu_is_stale(tape_)) return;
    if (ffgpu_set_prime(tape_, FFInt::p)) die(ffgpu_last_error());
    prime_internal = FFInt::p;
    warm_up_once();
  }
  // The first evaluation of a process loads the kernels it launches (lazy loading: 31 ms on a B200,
  // profiles/r02_first_call.log).  The first parser that has functions pays it here, so that it does not show up in
  // what the Reconstructor reports as the time of a black-box probe (see GpuBlackBox::warm_up).
  void warm_up_once() const {
    static std::atomic<bool> done{false};
    const size_t n = ffgpu_n_vars(tape_), m = ffgpu_n_out(tape_);
    if (m == 0 || !ctx() || done.exchange(true)) return;      // no GPU: a host-only parser (RPN, sizes) evaluates nothing
    std::vector<uint64_t> ones(std::max<size_t>(n, 1), 1), out(m);
    if (ffgpu_eval(tape_, ones.data(), 1, out.data(), 0)) die(ffgpu_last_error());
  }

  // --- evaluation ----------------------------------------------------------------------------------------------
  // Re-entrant: called concurrently from the Reconstructor's thread pool (Reconstructor.hpp:2981-2983); each
  // calling thread is mapped to a stream slot of its own.
  template <typename FFIntTemp>
  std::vector<FFIntTemp> evaluate_pre(const std::vector<FFIntTemp>& values) const noexcept {
    static_assert(sizeof(FFInt) == sizeof(uint64_t), "FFInt must be one 64-bit residue");
    static_assert(sizeof(FFIntTemp) % sizeof(uint64_t) == 0, "bunch type must be an array of FFInt");
    constexpr size_t lanes = sizeof(FFIntTemp) / sizeof(uint64_t);      // 1 for FFInt, N for FFIntVec<N>
    if (!tape_) return std::vector<FFIntTemp>();
    if (values.size() < ffgpu_n_vars(tape_)) die("parser called with too few variables");
    std::vector<FFIntTemp> result(ffgpu_n_out(tape_));
    const uint32_t slot = gpu_caller_slot();                   // one caller slot per thread, in order of first appearance
    if (ffgpu_eval(tape_, reinterpret_cast<const uint64_t*>(values.data()), lanes,
                   reinterpret_cast<uint64_t*>(result.data()), slot))
      die(ffgpu_last_error());
    return result;
  }
  template <typename FFIntTemp>
  std::vector<FFIntTemp> evaluate(const std::vector<FFIntTemp>& values) const { return evaluate_pre(values); }
  template <typename FFIntTemp>
  std::vector<FFIntTemp> evaluate_pre_2(const std::vector<FFIntTemp>& values) const noexcept { return evaluate_pre(values); }

  // --- inspection ----------------------------------------------------------------------------------------------
  std::vector<std::vector<std::string>> get_rp_functions() const { return rpn(); }
  void move_rpn(std::vector<std::vector<std::string>>& rpn_) { rpn_ = rpn(); }
  std::vector<std::string> get_rp_function(size_t i) const { return rpn().at(i); }
  const std::vector<std::vector<std::string>>* get_rp_functions_ref() const { return &rpn(); }
  bool empty() const { return get_size() == 0; }
  size_t get_size() const { return tape_ ? ffgpu_n_functions(tape_) : 0; }
  // number of values evaluate_pre returns (differs from get_size() on a check_is_equal parser)
  size_t get_n_out() const { return tape_ ? ffgpu_n_out(tape_) : 0; }
  ffgpu_tape* handle() { return tape_; }

 private:
  static std::vector<int>& device_list() {
    static std::vector<int> d;
    return d;
  }
  // one context per process, alive until exit; NULL without a GPU (host-only parsers, evaluation refuses)
  static ffgpu_ctx* ctx() {
    static ffgpu_ctx* c = [] {
      const std::vector<int>& d = device_list();
      return ffgpu_create(d.empty() ? nullptr : d.data(), static_cast<int>(d.size()));
    }();
    return c;
  }
  static std::vector<const char*> c_strs(const std::vector<std::string>& s) {
    std::vector<const char*> v;
    v.reserve(s.size());
    for (const auto& x : s) v.push_back(x.c_str());
    return v;
  }
  [[noreturn]] static void die(const std::string& msg) {
    ERROR_MSG(msg);
    std::exit(EXIT_FAILURE);
  }
  unsigned flags() const { return check_is_equal ? static_cast<unsigned>(FFGPU_CHECK_IS_EQUAL) : 0u; }
  void need_tape(const char* what) const {
    if (!tape_) die(std::string(what) + " on a parser without variables: construct it with vars or call parse_function first");
  }
  const std::vector<std::vector<std::string>>& rpn() const {
    if (!rpn_valid_) {
      rpn_cache_.clear();
      const size_t n = get_size();
      rpn_cache_.resize(n);
      std::string buf;
      for (size_t i = 0; i < n; ++i) {
        buf.resize(ffgpu_rpn(tape_, i, nullptr, 0) + 1);
        ffgpu_rpn(tape_, i, &buf[0], buf.size());
        std::istringstream is(buf.c_str());
        for (std::string tok; is >> tok;) rpn_cache_[i].push_back(tok);
      }
      rpn_valid_ = true;
    }
    return rpn_cache_;
  }
  void swap(GpuShuntingYardParser& o) noexcept {
    std::swap(tape_, o.tape_);
    vars_.swap(o.vars_);
    std::swap(check_is_equal, o.check_is_equal);
    std::swap(keep_rpn, o.keep_rpn);
    std::swap(prime_internal, o.prime_internal);
    rpn_cache_.swap(o.rpn_cache_);
    std::swap(rpn_valid_, o.rpn_valid_);
  }

  ffgpu_tape* tape_ = nullptr;
  std::vector<std::string> vars_;
  bool check_is_equal = false;
  bool keep_rpn = false;            // accepted for source compatibility; the RPN is always kept
  uint64_t prime_internal = 0;
  mutable std::vector<std::vector<std::string>> rpn_cache_;
  mutable bool rpn_valid_ = false;
};

}  // namespace firefly
