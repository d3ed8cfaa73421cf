// queue.cpp -- the probe queue in front of the GPU evaluator (C ABI: ffgpu_queue_*, include/ffgpu.h).
//
// What it stands in for.  Reconstructor keeps two queues around the black box (reference:
// source/include/firefly/Reconstructor.hpp): `requested_probes`, a deque of (index, values) filled by
// queue_probes (:245, :2962-2971), and `computed_probes`, a queue of (indices, [function][lane]) drained by
// get_probe (:246, :2996-3012).  Between them thr_n pool threads run get_job (:3326-3477): pop
// compute_bunch_size(queue length, thr_n, bunch_size) <= 128 points (utils.cpp:305-330), transpose them to
// [variable][lane], call bb.eval, push the result.
//
// A GPU wants the opposite bunch policy: ONE consumer that takes everything that is waiting (up to
// max_batch) and evaluates it in one launch, so that batches grow with the backlog instead of being split
// over threads.  This file is that consumer: a worker thread, the same two queues, the same index
// bookkeeping (every probe gets the running index `ind`, results come back tagged with it, in any order).
// Probes can be queued as explicit points or -- what queue_probes really generates -- as points on a line
// x_v = a_v*t + s_v (:2946-2960), in which case only t travels to the GPU (ffgpu_eval_line).
//
// Host-only C++; evaluation goes through ffgpu_eval / ffgpu_eval_line (no CPU path).
#include <algorithm>
#include <condition_variable>
#include <cstring>
#include <deque>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/ffgpu.h"

namespace {

// The worker's caller slot (ffgpu_eval's thread_id): the last of the library's 32.  Pool threads are numbered in
// order of first appearance from 0 (include/firefly_gpu/caller_slot.hpp), so a queue and a pool of up to 31 threads
// in one process never wait for each other's slot.
constexpr uint32_t QUEUE_SLOT = 31;

struct Request {
  uint64_t first_index = 0;
  size_t n = 0;
  bool line = false;
  std::vector<uint64_t> data;   // explicit: n * n_vars values, point-major (AoS like requested_probes)
                                // line: a[n_vars], s[n_vars], t[n]
};
struct Result {
  uint64_t index;
  std::vector<uint64_t> values;   // n_out
};

}  // namespace

struct ffgpu_queue {
  ffgpu_tape* tape = nullptr;
  size_t n_vars = 0, n_out = 0, max_batch = 0;
  std::mutex mtx;
  std::condition_variable cv_work, cv_done;
  std::deque<Request> requested;      // requested_probes
  std::deque<Result> computed;        // computed_probes
  uint64_t next_index = 0;            // Reconstructor's `ind`
  size_t in_flight = 0;               // probes taken by the worker, not yet in `computed`
  size_t queued = 0;                  // probes in `requested`
  uint64_t batches = 0, probes_done = 0, largest_batch = 0;
  bool stop = false;
  std::string error;
  std::thread worker;
  // pinned staging, owned by the worker
  uint64_t* h_in = nullptr;
  uint64_t* h_out = nullptr;
  size_t cap_in = 0, cap_out = 0;

  bool ensure(size_t need_in, size_t need_out) {
    if (cap_in < need_in) {
      ffgpu_host_free(h_in);
      h_in = static_cast<uint64_t*>(ffgpu_host_alloc(need_in * sizeof(uint64_t)));
      cap_in = h_in ? need_in : 0;
    }
    if (cap_out < need_out) {
      ffgpu_host_free(h_out);
      h_out = static_cast<uint64_t*>(ffgpu_host_alloc(need_out * sizeof(uint64_t)));
      cap_out = h_out ? need_out : 0;
    }
    return h_in && h_out;
  }

  void run() {
    std::vector<Request> batch;
    for (;;) {
      batch.clear();
      size_t n_explicit = 0;
      {
        std::unique_lock<std::mutex> lk(mtx);
        cv_work.wait(lk, [&] { return stop || !requested.empty(); });
        if (stop) return;
        // the GPU bunch policy: everything that waits, up to max_batch; line requests go one at a time
        // (each is already a batch), explicit points are merged
        if (requested.front().line) {
          batch.push_back(std::move(requested.front()));
          requested.pop_front();
        } else {
          while (!requested.empty() && !requested.front().line && n_explicit + requested.front().n <= max_batch) {
            n_explicit += requested.front().n;
            batch.push_back(std::move(requested.front()));
            requested.pop_front();
          }
          if (batch.empty()) {      // a single request larger than max_batch
            n_explicit = requested.front().n;
            batch.push_back(std::move(requested.front()));
            requested.pop_front();
          }
        }
        size_t taken = 0;
        for (const Request& r : batch) taken += r.n;
        queued -= taken;
        in_flight += taken;
      }
      size_t n = 0;
      for (const Request& r : batch) n += r.n;
      std::string err;
      // the tape may have been rebuilt since the queue was created (ffgpu_add_function / ffgpu_trim /
      // ffgpu_set_factors + ffgpu_set_prime): requests and staging are laid out for the shape seen at creation
      const bool same_shape = ffgpu_n_vars(tape) == n_vars && ffgpu_n_out(tape) == n_out;
      if (!same_shape) err = "the collection changed its number of variables or outputs while a queue was attached";
      bool ok = same_shape && ensure(std::max<size_t>(n * std::max<size_t>(n_vars, 1), 1), std::max<size_t>(n * std::max<size_t>(n_out, 1), 1));
      if (!ok && err.empty()) err = "pinned staging allocation failed";
      if (ok) {
        if (batch[0].line) {
          const Request& r = batch[0];
          const uint64_t* a = r.data.data();
          const uint64_t* s = a + n_vars;
          std::memcpy(h_in, s + n_vars, n * sizeof(uint64_t));
          ok = ffgpu_eval_line(tape, a, s, h_in, n, h_out, QUEUE_SLOT) == 0;
        } else {
          // AoS -> SoA, what compute_probe<N> does for its FFIntVec (Reconstructor.hpp:3424-3434)
          size_t j = 0;
          for (const Request& r : batch)
            for (size_t q = 0; q < r.n; ++q, ++j)
              for (size_t v = 0; v < n_vars; ++v) h_in[v * n + j] = r.data[q * n_vars + v];
          ok = ffgpu_eval(tape, h_in, n, h_out, QUEUE_SLOT) == 0;
        }
        if (!ok) err = ffgpu_last_error();
      }
      // results are unpacked before the lock is taken: producers and consumers only wait for the hand-over
      std::vector<Result> done;
      if (ok) {
        done.reserve(n);
        size_t j = 0;
        for (const Request& r : batch)
          for (size_t q = 0; q < r.n; ++q, ++j) {
            Result res;
            res.index = r.first_index + q;
            res.values.resize(n_out);
            for (size_t f = 0; f < n_out; ++f) res.values[f] = h_out[f * n + j];
            done.push_back(std::move(res));
          }
      }
      {
        std::lock_guard<std::mutex> lk(mtx);
        if (ok) {
          for (Result& res : done) computed.push_back(std::move(res));
          ++batches;
          probes_done += n;
          if (n > largest_batch) largest_batch = n;
        } else if (error.empty()) {
          error = err;
        }
        in_flight -= n;
      }
      cv_done.notify_all();
    }
  }
};

extern "C" {

// MPIWorker's wire format in one call (reference: source/include/firefly/MPIWorker.hpp:213-228 receives
// VALUES = [index, x_1 .. x_n] * k, :479-501 answers RESULT = [index, f_1 .. f_m] * k): transpose the message to
// [variable][probe], evaluate all k probes in one batch, interleave the answers.  A GPU node can serve an
// unmodified FireFly MPI master with this and two MPI calls.
int ffgpu_eval_wire(ffgpu_tape* tape, const uint64_t* values_msg, size_t k, uint64_t* result_msg, uint32_t thread_id) {
  if (!tape) return ffgpu_eval(nullptr, nullptr, 0, nullptr, thread_id);   // sets ffgpu_last_error ("null tape")
  if (k == 0) return 0;
  const size_t n = ffgpu_n_vars(tape), m = ffgpu_n_out(tape);
  std::vector<uint64_t> in(std::max<size_t>(n * k, 1)), out(std::max<size_t>(m * k, 1));
  for (size_t j = 0; j < k; ++j)
    for (size_t v = 0; v < n; ++v) in[v * k + j] = values_msg[j * (n + 1) + 1 + v];
  const int rc = ffgpu_eval(tape, in.data(), k, out.data(), thread_id);
  if (rc) return rc;
  for (size_t j = 0; j < k; ++j) {
    result_msg[j * (m + 1)] = values_msg[j * (n + 1)];
    for (size_t f = 0; f < m; ++f) result_msg[j * (m + 1) + 1 + f] = out[f * k + j];
  }
  return 0;
}

ffgpu_queue* ffgpu_queue_create(ffgpu_tape* tape, size_t max_batch) {
  if (!tape) return nullptr;
  ffgpu_queue* q = new ffgpu_queue;
  q->tape = tape;
  q->n_vars = ffgpu_n_vars(tape);
  q->n_out = ffgpu_n_out(tape);
  q->max_batch = max_batch ? max_batch : (size_t)1 << 16;
  q->worker = std::thread([q] { q->run(); });
  return q;
}

void ffgpu_queue_destroy(ffgpu_queue* q) {
  if (!q) return;
  {
    std::lock_guard<std::mutex> lk(q->mtx);
    q->stop = true;
  }
  q->cv_work.notify_all();
  if (q->worker.joinable()) q->worker.join();
  ffgpu_host_free(q->h_in);
  ffgpu_host_free(q->h_out);
  delete q;
}

uint64_t ffgpu_queue_push(ffgpu_queue* q, const uint64_t* values) {
  Request r;
  r.n = 1;
  r.data.assign(values, values + q->n_vars);
  uint64_t index;
  {
    std::lock_guard<std::mutex> lk(q->mtx);
    index = r.first_index = q->next_index++;
    q->queued += 1;
    q->requested.push_back(std::move(r));
  }
  q->cv_work.notify_one();
  return index;
}

uint64_t ffgpu_queue_push_many(ffgpu_queue* q, const uint64_t* values, size_t n) {
  Request r;
  r.n = n;
  r.data.assign(values, values + n * q->n_vars);
  uint64_t first;
  {
    std::lock_guard<std::mutex> lk(q->mtx);
    first = r.first_index = q->next_index;
    q->next_index += n;
    q->queued += n;
    if (n) q->requested.push_back(std::move(r));
  }
  q->cv_work.notify_one();
  return first;
}

uint64_t ffgpu_queue_push_line(ffgpu_queue* q, const uint64_t* a, const uint64_t* s, const uint64_t* t, size_t n) {
  Request r;
  r.n = n;
  r.line = true;
  r.data.reserve(2 * q->n_vars + n);
  r.data.insert(r.data.end(), a, a + q->n_vars);
  r.data.insert(r.data.end(), s, s + q->n_vars);
  r.data.insert(r.data.end(), t, t + n);
  uint64_t first;
  {
    std::lock_guard<std::mutex> lk(q->mtx);
    first = r.first_index = q->next_index;
    q->next_index += n;
    q->queued += n;
    if (n) q->requested.push_back(std::move(r));
  }
  q->cv_work.notify_one();
  return first;
}

long ffgpu_queue_pop(ffgpu_queue* q, uint64_t* indices, uint64_t* out, size_t cap, int wait) {
  std::unique_lock<std::mutex> lk(q->mtx);
  if (wait)
    q->cv_done.wait(lk, [&] { return !q->computed.empty() || !q->error.empty() || (q->queued + q->in_flight == 0); });
  if (!q->error.empty()) return -1;
  size_t n = 0;
  while (n < cap && !q->computed.empty()) {
    const Result& r = q->computed.front();
    indices[n] = r.index;
    for (size_t f = 0; f < q->n_out; ++f) out[f * cap + n] = r.values[f];
    q->computed.pop_front();
    ++n;
  }
  return (long)n;
}

size_t ffgpu_queue_pending(ffgpu_queue* q) {
  std::lock_guard<std::mutex> lk(q->mtx);
  return q->queued + q->in_flight;
}

void ffgpu_queue_stats(ffgpu_queue* q, uint64_t s[3]) {
  std::lock_guard<std::mutex> lk(q->mtx);
  s[0] = q->batches;
  s[1] = q->probes_done;
  s[2] = q->largest_batch;
}

const char* ffgpu_queue_error(ffgpu_queue* q) {
  std::lock_guard<std::mutex> lk(q->mtx);
  return q->error.c_str();
}

}  // extern "C"
