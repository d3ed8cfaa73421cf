// GpuProbeQueue.hpp -- header-only C++ face of ffgpu_queue_* (include/ffgpu.h): the pair of queues a
// Reconstructor keeps around its black box, with the GPU as the consumer.
//
// Reference shape (source/include/firefly/Reconstructor.hpp): queue_probes() appends (index, values) to
// `requested_probes` (:2962-2971) and spawns get_job tasks; get_probe() pops (indices, [function][lane]) from
// `computed_probes` (:2996-3012).  Here push()/push_line() are the producer side and pop() the consumer side;
// the get_job threads in between are replaced by one worker inside the library that evaluates the whole
// backlog per launch.  push_line() takes what queue_probes really generates -- the parameters t of points
// x_v = a_v * t + s_v for one zi_order (:2946-2960) -- so only t crosses PCIe.
//
// Plain uint64_t residues: FFInt is one uint64_t (FFInt.hpp:89), so std::vector<FFInt> data can be passed
// through reinterpret_cast like in GpuBlackBox.hpp.  No FireFly header is needed to use this one.
#pragma once
#include <cstdint>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "../ffgpu.h"

namespace firefly {

class GpuProbeQueue {
 public:
  // `tape` stays owned by the caller (e.g. GpuBlackBox::handle()) and must outlive the queue
  explicit GpuProbeQueue(ffgpu_tape* tape, size_t max_batch = 0)
      : q_(ffgpu_queue_create(tape, max_batch)), n_vars_(ffgpu_n_vars(tape)), n_out_(ffgpu_n_out(tape)) {
    if (!q_) throw std::runtime_error("GpuProbeQueue: could not create the queue");
  }
  GpuProbeQueue(const GpuProbeQueue&) = delete;
  GpuProbeQueue& operator=(const GpuProbeQueue&) = delete;
  ~GpuProbeQueue() { ffgpu_queue_destroy(q_); }

  // one probe point; returns its index (Reconstructor's `ind`)
  uint64_t push(const std::vector<uint64_t>& values) {
    if (values.size() != n_vars_) throw std::invalid_argument("GpuProbeQueue::push: wrong number of variables");
    return ffgpu_queue_push(q_, values.data());
  }
  // n points on the line x_v = a[v] * t[j] + s[v]; returns the index of the first
  uint64_t push_line(const std::vector<uint64_t>& a, const std::vector<uint64_t>& s, const std::vector<uint64_t>& t) {
    if (a.size() != n_vars_ || s.size() != n_vars_) throw std::invalid_argument("GpuProbeQueue::push_line: wrong number of variables");
    return ffgpu_queue_push_line(q_, a.data(), s.data(), t.data(), t.size());
  }
  // up to `cap` finished probes as (indices, values[function][probe]); blocks while probes are outstanding and
  // none is ready; both empty when nothing is outstanding
  std::pair<std::vector<uint64_t>, std::vector<std::vector<uint64_t>>> pop(size_t cap = 4096, bool wait = true) {
    std::vector<uint64_t> idx(cap), flat(cap * (n_out_ ? n_out_ : 1));
    const long k = ffgpu_queue_pop(q_, idx.data(), flat.data(), cap, wait ? 1 : 0);
    if (k < 0) throw std::runtime_error(std::string("GpuProbeQueue: ") + ffgpu_queue_error(q_));
    idx.resize((size_t)k);
    std::vector<std::vector<uint64_t>> vals(n_out_, std::vector<uint64_t>((size_t)k));
    for (size_t f = 0; f < n_out_; ++f)
      for (size_t j = 0; j < (size_t)k; ++j) vals[f][j] = flat[f * cap + j];
    return {std::move(idx), std::move(vals)};
  }
  size_t pending() const { return ffgpu_queue_pending(q_); }

 private:
  ffgpu_queue* q_;
  size_t n_vars_, n_out_;
};

}  // namespace firefly
