// GpuBlackBox.hpp -- header-only C++ adaptor: a FireFly black box whose evaluation runs on B200s.
//
// Drop-in for the parser-backed black boxes FireFly users write, e.g. the one in the reference's
// benchmarks/benchmarks.cpp:22-38 or AmplitudeParser.hpp:45-68 (FFAmplitudeBlackBox):
//
//     class BlackBoxUser : public BlackBoxBase<BlackBoxUser> {
//       template<typename FFIntTemp> std::vector<FFIntTemp> operator()(const std::vector<FFIntTemp>& values)
//         { return par.evaluate_pre(values); }
//       void prime_changed() { par.precompute_tokens(); }
//       ShuntingYardParser par;
//     };
//
// becomes   firefly::GpuBlackBox bb(file, vars);   Reconstructor<GpuBlackBox> reconst(n, threads, bunch, bb);
//
// Keeps the 2.0.3 plugin boundary (reference: source/include/firefly/BlackBoxBase.hpp:35-45):
// operator()(values) for FFInt and every FFIntVec<N>, prime_changed().  Adds the forward-compatible
// operator()(values, thread_id) and set_bunch_size() named by newer FireFly interfaces.
//
// Zero-copy at the boundary: FFInt is a single uint64_t (FFInt.hpp:89) and FFIntVec<N> a
// std::array<FFInt, N> (FFIntVec.hpp:32-109), so std::vector<FFIntVec<N>>(n) IS the [variable][lane]
// uint64 matrix ffgpu_eval wants.
//
// Errors follow the reference's convention for this path: ERROR_MSG + std::exit(EXIT_FAILURE)
// (ShuntingYardParser.hpp:454-459).
#pragma once
#include <algorithm>
#include <cstdint>
#include <cstdlib>
#include <functional>
#include <string>
#include <atomic>
#include <thread>
#include <vector>

#include "firefly/BlackBoxBase.hpp"
#include "firefly/FFInt.hpp"
#include "firefly/FFIntVec.hpp"
#include "firefly/Logger.hpp"

#include "../ffgpu.h"
#include "caller_slot.hpp"

namespace firefly {

class GpuBlackBox : public BlackBoxBase<GpuBlackBox> {
 public:
  // functions in a ';'-separated file, like ShuntingYardParser(file, vars, check_is_equal)
  GpuBlackBox(const std::string& file, const std::vector<std::string>& vars, bool check_is_equal = false,
              const std::vector<int>& devices = {}) {
    init(devices);
    std::vector<const char*> v = c_strs(vars);
    tape_ = ffgpu_compile_file(ctx_, file.c_str(), v.data(), v.size(), check_is_equal ? FFGPU_CHECK_IS_EQUAL : 0);
    finish();
  }
  // functions as strings, like ShuntingYardParser(funs, vars, check_is_equal)
  GpuBlackBox(const std::vector<std::string>& funs, const std::vector<std::string>& vars, bool check_is_equal = false,
              const std::vector<int>& devices = {}) {
    init(devices);
    std::vector<const char*> f = c_strs(funs), v = c_strs(vars);
    tape_ = ffgpu_compile(ctx_, f.data(), f.size(), v.data(), v.size(), check_is_equal ? FFGPU_CHECK_IS_EQUAL : 0);
    finish();
  }
  GpuBlackBox(const GpuBlackBox&) = delete;
  GpuBlackBox& operator=(const GpuBlackBox&) = delete;
  ~GpuBlackBox() {
    if (tape_) ffgpu_tape_destroy(tape_);
    if (ctx_) ffgpu_destroy(ctx_);
  }

  // FireFly 2.0.3 signature.  Called concurrently from the Reconstructor's thread pool
  // (Reconstructor.hpp:2981-2983); each calling thread is mapped to its own stream slot.
  template <typename FFIntTemp>
  std::vector<FFIntTemp> operator()(const std::vector<FFIntTemp>& values) {
    return (*this)(values, caller_slot());
  }

  // Every calling thread gets a number of its own in order of first appearance: the library maps it to one of its
  // caller slots (stream + staging), so a pool of up to 32 threads never shares one (a hash of the thread id did).
  static uint32_t caller_slot() { return gpu_caller_slot(); }

  template <typename FFIntTemp>
  std::vector<FFIntTemp> operator()(const std::vector<FFIntTemp>& values, uint32_t thread_id) {
    static_assert(sizeof(FFInt) == sizeof(uint64_t), "FFInt must be one 64-bit residue");
    static_assert(sizeof(FFIntTemp) % sizeof(uint64_t) == 0, "bunch type must be an array of FFInt");
    constexpr size_t lanes = sizeof(FFIntTemp) / sizeof(uint64_t);      // 1 for FFInt, N for FFIntVec<N>
    if (values.size() != n_vars_) die("black box called with a wrong number of variables");
    std::vector<FFIntTemp> result(n_out_);
    if (ffgpu_eval(tape_, reinterpret_cast<const uint64_t*>(values.data()), lanes,
                   reinterpret_cast<uint64_t*>(result.data()), thread_id))
      die(ffgpu_last_error());
    return result;
  }

  // Reconstructor calls this after FFInt::set_new_prime (Reconstructor.hpp:2483-2484), never
  // concurrently with operator().  Same job as par.precompute_tokens(): refresh the constants.
  void prime_changed() {
    if (ffgpu_set_prime(tape_, FFInt::p)) die(ffgpu_last_error());
  }

  // The bunch size is a Reconstructor constructor argument in 2.0.3 (Reconstructor.hpp:73): every pool thread calls
  // with up to `bunch` points.  The GPU evaluator accepts any batch.  If the aggregator of concurrent small calls is
  // switched on (set_coalesce(true) or FFGPU_COALESCE=1; off by default, see ffgpu.h), the bunch size sets how many
  // points a shared launch may hold -- room for 32 callers' bunches, at least 512 points.
  void set_bunch_size(uint32_t bunch) {
    bunch_hint_ = bunch ? bunch : 1;
    if (ffgpu_set_coalesce(tape_, coalesce_, std::max<size_t>(512, (size_t)32 * bunch_hint_))) die(ffgpu_last_error());
  }
  // concurrent small calls share launches (fewer launches, more latency per call)
  void set_coalesce(bool on) {
    coalesce_ = on ? 1 : 0;
    set_bunch_size(bunch_hint_);
  }
  uint32_t bunch_size() const { return bunch_hint_; }
  // launches that the aggregator shared among callers so far, calls they served, points
  void coalesce_stats(uint64_t stats[3]) const { ffgpu_coalesce_stats(tape_, stats); }

  size_t n_vars() const { return n_vars_; }
  size_t n_out() const { return n_out_; }
  ffgpu_tape* handle() { return tape_; }

 private:
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
  void init(const std::vector<int>& devices) {
    ctx_ = ffgpu_create(devices.empty() ? nullptr : devices.data(), (int)devices.size());
    if (!ctx_) die(ffgpu_last_error());
  }
  void finish() {
    if (!tape_) die(ffgpu_last_error());
    n_vars_ = ffgpu_n_vars(tape_);
    n_out_ = ffgpu_n_out(tape_);
    prime_changed();      // ShuntingYardParser's constructor precomputes for the current prime (:181)
    warm_up();
  }
  // The first evaluation of a process loads the kernels it launches (the CUDA runtime loads them lazily: 31 ms
  // measured on a B200, profiles/r02_first_call.log) and sets up the caller slot.  Paid here, once, it stays out of
  // what the Reconstructor reports as the time of a black-box probe: every pool thread's first call would wait for it.
  void warm_up() {
    if (!n_vars_ && !n_out_) return;
    for (size_t lanes : {(size_t)1, (size_t)128}) {
      std::vector<uint64_t> ones(std::max<size_t>(n_vars_ * lanes, 1), 1), out(std::max<size_t>(n_out_ * lanes, 1));
      if (ffgpu_eval(tape_, ones.data(), lanes, out.data(), 0)) die(ffgpu_last_error());
    }
  }

  ffgpu_ctx* ctx_ = nullptr;
  ffgpu_tape* tape_ = nullptr;
  size_t n_vars_ = 0, n_out_ = 0;
  uint32_t bunch_hint_ = 1;
  int coalesce_ = -1;          // -1: environment FFGPU_COALESCE / default (off)
};

}  // namespace firefly
