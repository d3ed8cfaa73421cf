// capi.cu -- the C ABI of include/ffgpu.h: contexts, compiled tapes, host-buffer evaluation
// with multi-GPU sharding and copy/compute pipelining, device-buffer evaluation.
//
// Host-side counterpart of what Reconstructor::get_job / compute_probe<N> do around bb.eval
// (reference: source/include/firefly/Reconstructor.hpp:3325-3477): take a batch of probe points,
// run the black box, hand back [function][point].  Here the batch is sharded over the GPUs of
// the context by contiguous point ranges (no communication between GPUs -- probes are
// independent), and every shard is cut into chunks whose H2D copy, kernel and D2H copy run on
// four rotating streams so that copies overlap compute.
#include <cuda_runtime.h>

#include <atomic>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <memory>
#include <mutex>
#include <sstream>
#include <string>
#include <thread>
#include <vector>

#include "../../include/ffgpu.h"
#include "compiler.hpp"
#include "kernels.cuh"

namespace {

thread_local std::string g_err;
std::atomic<uint64_t> g_launches{0};

int fail(const std::string& msg) {
  g_err = msg;
  return -1;
}
int cuda_fail(cudaError_t e, const char* what) {
  g_err = std::string(what) + ": " + cudaGetErrorString(e);
  return -1;
}
#define FF_CUDA(call)                                   \
  do {                                                  \
    cudaError_t e_ = (call);                            \
    if (e_ != cudaSuccess) return cuda_fail(e_, #call); \
  } while (0)

constexpr int N_SLOTS = 32;       // concurrent ffgpu_eval callers served without contention
constexpr int N_BUF = 4;          // chunks in flight per slot and device (H2D / kernels / D2H of different chunks overlap)

struct DevBuf {
  cudaStream_t stream = nullptr;
  uint64_t* d_in = nullptr;
  uint64_t* d_out = nullptr;
  uint64_t* d_as = nullptr;         // line mode: a[n_vars] then s[n_vars]
  uint64_t* h_as = nullptr;         // pinned copy of a, s: uploaded once per call and buffer, asynchronously
  uint64_t* h_stage = nullptr;      // pinned: small calls travel as one contiguous copy each way
  size_t cap_in = 0, cap_out = 0, cap_as = 0, cap_stage = 0;   // elements
};
struct SlotDev {
  DevBuf buf[N_BUF];
};
struct Slot {
  std::mutex mtx;
  std::vector<SlotDev> dev;
};
struct Device {
  int id = 0;
  int sm_count = 148;
};

// ---- the 300 largest primes below 2^63 (what the reference tabulates in ReconstHelper.cpp:24-125),
// found by a deterministic Miller-Rabin search instead of a table
typedef unsigned __int128 u128;
uint64_t mulmod64(uint64_t a, uint64_t b, uint64_t m) { return (uint64_t)((u128)a * b % m); }
uint64_t powmod64(uint64_t b, uint64_t e, uint64_t m) {
  uint64_t r = 1;
  for (; e; e >>= 1, b = mulmod64(b, b, m))
    if (e & 1) r = mulmod64(r, b, m);
  return r;
}
bool is_prime64(uint64_t n) {
  if (n < 2) return false;
  for (uint64_t q : {2ULL, 3ULL, 5ULL, 7ULL, 11ULL, 13ULL, 17ULL, 19ULL, 23ULL, 29ULL, 31ULL, 37ULL}) {
    if (n % q == 0) return n == q;
  }
  uint64_t d = n - 1;
  int s = 0;
  while (!(d & 1)) { d >>= 1; ++s; }
  for (uint64_t a : {2ULL, 3ULL, 5ULL, 7ULL, 11ULL, 13ULL, 17ULL, 19ULL, 23ULL, 29ULL, 31ULL, 37ULL}) {
    uint64_t x = powmod64(a, d, n);
    if (x == 1 || x == n - 1) continue;
    bool comp = true;
    for (int i = 1; i < s && comp; ++i) {
      x = mulmod64(x, x, n);
      if (x == n - 1) comp = false;
    }
    if (comp) return false;
  }
  return true;
}
const std::vector<uint64_t>& prime_table() {
  static const std::vector<uint64_t> tab = [] {
    std::vector<uint64_t> t;
    for (uint64_t n = (1ULL << 63) - 1; t.size() < 300; n -= 2)
      if (is_prime64(n)) t.push_back(n);
    return t;
  }();
  return tab;
}

}  // namespace

struct ffgpu_ctx {
  std::vector<Device> devs;
  Slot slots[N_SLOTS];
};

// Aggregator of concurrent small calls (ffgpu_eval with a few points each, from many threads -- what Reconstructor's
// thread pool does: get_job hands every pool thread a bunch of <= 128 points, reference: Reconstructor.hpp:3325-3477,
// utils.cpp:305-330).  Callers deposit their points into the batch that is OPEN; the first one in becomes its leader,
// waits until the batch before it has left the GPU -- every call that arrives meanwhile joins --, closes it, runs
// ONE launch for everybody and wakes the others; each caller takes its slice.  A caller that is alone is its own
// leader at once: nothing waits for company.  Two batches alternate (one fills while the other runs), each with
// pinned, mapped staging and device buffers of its own; on a context with several GPUs the two use different ones.
struct CoBatch {
  uint64_t *h_in = nullptr, *h_out = nullptr, *d_in = nullptr;      // h_*: one pinned mapped allocation (h_in)
  cudaStream_t stream = nullptr;
  size_t cap = 0, n_vars = 0, n_out = 0;     // layout [row][cap] the buffers were allocated for
  int dev = -1;                              // index into ctx->devs of d_in / stream
  size_t count = 0;                          // points deposited
  unsigned users = 0;                        // callers that still have to take their slice
  bool has_leader = false, closed = false, done = false;
  uint64_t seq = 0;
  std::atomic<uint64_t> done_seq{0};         // = seq once the batch's results are in h_out (followers spin on it)
  int rc = 0;
  std::string err;
};
struct Coalescer {
  std::mutex m;
  std::condition_variable cv;
  CoBatch b[2];
  int open = 0;
  bool busy = false;                         // a batch is on the GPU
  uint64_t next_seq = 1, batches = 0, calls = 0, points = 0;
};

struct ffgpu_tape {
  ffgpu_ctx* ctx = nullptr;
  std::unique_ptr<Coalescer> co{new Coalescer};
  int coalesce = -1;                   // -1: FFGPU_COALESCE / default off
  size_t coalesce_cap = 0;             // points per batch (0: automatic)
  ffgpu::Collection col;
  std::vector<ffgpu::Instr*> d_tape;   // per device
  std::vector<uint32_t*> d_chunks;     // per device: Collection::chunk_table
  // per level of shared rows (Collection::pre: 0 = common sub-expressions, 1 = partial sums) and device: tape, chunk table
  std::vector<ffgpu::Instr*> d_pre[2];
  std::vector<uint32_t*> d_pre_chunks[2];
  bool pre_needs_generic[2] = {true, true};
  int force_k = 0, force_threads = 0;
  size_t chunk = 0;
  bool needs_generic = true;           // the resolved tape has instructions without a fast dispatch code
};

namespace {

struct DeviceGuard {
  int prev = 0;
  bool ok = false;
  explicit DeviceGuard(int dev) {
    ok = cudaGetDevice(&prev) == cudaSuccess && cudaSetDevice(dev) == cudaSuccess;
  }
  ~DeviceGuard() { cudaSetDevice(prev); }
};

ffgpu_tape* finish_compile(ffgpu_ctx* ctx, std::unique_ptr<ffgpu_tape> t, bool ok, const std::string& err) {
  if (!ok) {
    g_err = err;
    return nullptr;
  }
  t->ctx = ctx;       // NULL: host-only handle (RPN / tape inspection; evaluation refuses)
  if (ctx) {
    t->d_tape.assign(ctx->devs.size(), nullptr);
    t->d_chunks.assign(ctx->devs.size(), nullptr);
    for (int l = 0; l < 2; ++l) {
      t->d_pre[l].assign(ctx->devs.size(), nullptr);
      t->d_pre_chunks[l].assign(ctx->devs.size(), nullptr);
    }
  }
  return t.release();
}

std::vector<std::string> to_vec(const char* const* s, size_t n) {
  std::vector<std::string> v;
  v.reserve(n);
  for (size_t i = 0; i < n; ++i) v.emplace_back(s[i] ? s[i] : "");
  return v;
}

// one kernel launch: collection `c` (tape and chunk table on device g) at n_points points
int launch_one(ffgpu_tape* t, const ffgpu::Collection& c, const ffgpu::Instr* d_tape, const uint32_t* d_chunks,
               bool needs_generic, int g, const uint64_t* d_values, size_t in_stride, size_t n_points, uint64_t* d_out,
               size_t out_stride, const uint64_t* d_table, size_t table_stride, cudaStream_t stream) {
  ffgpu::EvalParams prm;
  prm.tape = d_tape;
  prm.values = d_values;
  prm.out = d_out;
  prm.n_points = n_points;
  prm.in_stride = in_stride;
  prm.out_stride = out_stride;
  prm.p = c.prime;
  prm.generic = ffgpu::is_table_form(c.prime) ? 0u : 1u;
  prm.c = prm.generic ? 0u : (uint32_t)((1ULL << 63) - c.prime);
  prm.d = 2 * prm.c;
  prm.pinv = 0;
  prm.pshift = 0;
  if (prm.generic) ffgpu::barrett_constants(c.prime, &prm.pinv, &prm.pshift);
  prm.n_slots = c.n_slots;
  prm.n_instr = (uint32_t)(c.tape.size() - 2);
  prm.n_vars = c.n_input_vars ? c.n_input_vars : (uint32_t)c.vars.size();     // a rows' collection also counts rows as variables
  prm.table = d_table;
  prm.n_base = d_table ? prm.n_vars : 0xffffffffu;
  prm.table_stride = table_stride;
  const ffgpu::LaunchShape sh = ffgpu::choose_shape(n_points, c.n_slots, t->ctx->devs[g].sm_count, t->force_k,
                                                    t->force_threads, c.max_chunks, !needs_generic, prm.generic != 0);
  prm.chunk_bounds = nullptr;
  ffgpu::LaunchShape shape = sh;
  static const bool no_exact = std::getenv("FFGPU_NO_EXACT_CHUNKS") != nullptr;       // A/B switch (development)
  if (!no_exact && c.exact_chunks > shape.chunks && shape.chunks == c.max_chunks && !t->force_k && !t->force_threads &&
      (uint64_t)shape.blocks * shape.chunks < 4ull * (uint64_t)t->ctx->devs[g].sm_count) {
    // few points and more functions than the largest power-of-two split: one function per chunk
    shape.chunks = c.exact_chunks;
    prm.chunk_bounds = d_chunks + c.chunk_table_off[c.exact_table];
  } else if (shape.chunks > 1) {
    unsigned idx = 0;
    while ((1u << idx) < shape.chunks) ++idx;
    prm.chunk_bounds = d_chunks + c.chunk_table_off[idx];
  }
  cudaError_t e = ffgpu::launch_eval(prm, shape, stream);
  if (e != cudaSuccess) return cuda_fail(e, "launch_eval");
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return 0;
}

// Evaluate the collection on device g.  A collection with shared rows (Collection::pre) runs two kernels per range of
// points: the rows' own tape writes them into a stream-ordered scratch allocation, the functions' tape reads them as
// extra variables.  The ranges are sized so that the rows stay L2 resident between the two.
int launch(ffgpu_tape* t, int g, const uint64_t* d_values, size_t in_stride, size_t n_points, uint64_t* d_out,
           size_t out_stride, cudaStream_t stream) {
  const ffgpu::Collection& c = t->col;
  if (c.pre.empty())
    return launch_one(t, c, t->d_tape[g], t->d_chunks[g], t->needs_generic, g, d_values, in_stride, n_points, d_out,
                      out_stride, nullptr, 0, stream);
  const size_t rows = c.n_table_rows;
  constexpr size_t BUDGET = 96u << 20;
  size_t sub = n_points;
  if (rows * n_points * 8 > BUDGET) sub = std::max<size_t>((BUDGET / (rows * 8)) & ~(size_t)2047, 16384);
  const size_t stride = (std::min(sub, n_points) + 1) & ~(size_t)1;
  uint64_t* tab = nullptr;
  FF_CUDA(cudaMallocAsync(&tab, rows * stride * 8, stream));
  int rc = 0;
  for (size_t off = 0; off < n_points && !rc; off += sub) {
    const size_t cnt = std::min(sub, n_points - off);
    // the rows level by level (a level reads the rows of the levels before it), then the functions
    for (size_t l = 0; l < c.pre.size() && !rc; ++l)
      rc = launch_one(t, c.pre[l], t->d_pre[l][g], t->d_pre_chunks[l][g], t->pre_needs_generic[l], g, d_values + off, in_stride,
                      cnt, tab + (size_t)c.pre_row_offset()[l] * stride, stride, l ? tab : nullptr, stride, stream);
    if (!rc)
      rc = launch_one(t, c, t->d_tape[g], t->d_chunks[g], t->needs_generic, g, d_values + off, in_stride, cnt,
                      d_out + off, out_stride, tab, stride, stream);
  }
  const cudaError_t e = cudaFreeAsync(tab, stream);
  if (!rc && e != cudaSuccess) return cuda_fail(e, "cudaFreeAsync");
  return rc;
}

// small-call path: how much of the transfer the kernel does itself (FFGPU_SMALL_ZERO_COPY = 0 | 1 | 2)
int small_zero_copy() {
  static const int mode = [] {
    const char* e = std::getenv("FFGPU_SMALL_ZERO_COPY");
    return e && *e ? std::atoi(e) : 1;
  }();
  return mode;
}

// the device copies of the tape and its chunk table; ffgpu_set_prime uploads new ones
void drop_device_copies(ffgpu_tape* t) {
  for (size_t g = 0; t->ctx && g < t->ctx->devs.size(); ++g) {
    DeviceGuard guard(t->ctx->devs[g].id);
    cudaDeviceSynchronize();
    if (t->d_tape[g]) cudaFree(t->d_tape[g]);
    if (t->d_chunks[g]) cudaFree(t->d_chunks[g]);
    for (int l = 0; l < 2; ++l) {
      if (t->d_pre[l][g]) cudaFree(t->d_pre[l][g]);
      if (t->d_pre_chunks[l][g]) cudaFree(t->d_pre_chunks[l][g]);
      t->d_pre[l][g] = nullptr;
      t->d_pre_chunks[l][g] = nullptr;
    }
    t->d_tape[g] = nullptr;
    t->d_chunks[g] = nullptr;
  }
}

// rows x cnt words between two row-major matrices.  One strided copy while both pitches are within the
// limit of cudaMemcpy2D (cudaDeviceProp::memPitch, 2^31 - 1 bytes); beyond it -- a single call with 2^28
// points or more -- one contiguous copy per row.
int copy_rows(uint64_t* dst, size_t dst_stride, const uint64_t* src, size_t src_stride, size_t cnt, size_t rows,
              cudaMemcpyKind kind, cudaStream_t stream) {
  constexpr size_t MAX_PITCH = 0x7fffffffULL;
  if (dst_stride * 8 <= MAX_PITCH && src_stride * 8 <= MAX_PITCH) {
    FF_CUDA(cudaMemcpy2DAsync(dst, dst_stride * 8, src, src_stride * 8, cnt * 8, rows, kind, stream));
    return 0;
  }
  for (size_t r = 0; r < rows; ++r)
    FF_CUDA(cudaMemcpyAsync(dst + r * dst_stride, src + r * src_stride, cnt * 8, kind, stream));
  return 0;
}

void free_batch(ffgpu_ctx* ctx, CoBatch& b) {
  if (b.dev >= 0 && ctx) {
    DeviceGuard guard(ctx->devs[b.dev].id);
    if (b.stream) {
      cudaStreamSynchronize(b.stream);
      cudaStreamDestroy(b.stream);
    }
    if (b.d_in) cudaFree(b.d_in);
    if (b.h_in) cudaFreeHost(b.h_in);
  }
  b.h_in = b.h_out = b.d_in = nullptr;
  b.stream = nullptr;
  b.cap = 0;
  b.dev = -1;
}

int ensure(DevBuf& b, size_t need_in, size_t need_out) {
  if (!b.stream) FF_CUDA(cudaStreamCreateWithFlags(&b.stream, cudaStreamNonBlocking));
  if (b.cap_in < need_in) {
    FF_CUDA(cudaStreamSynchronize(b.stream));
    if (b.d_in) FF_CUDA(cudaFree(b.d_in));
    b.d_in = nullptr;
    b.cap_in = 0;
    FF_CUDA(cudaMalloc(&b.d_in, need_in * sizeof(uint64_t)));
    b.cap_in = need_in;
  }
  if (b.cap_out < need_out) {
    FF_CUDA(cudaStreamSynchronize(b.stream));
    if (b.d_out) FF_CUDA(cudaFree(b.d_out));
    b.d_out = nullptr;
    b.cap_out = 0;
    FF_CUDA(cudaMalloc(&b.d_out, need_out * sizeof(uint64_t)));
    b.cap_out = need_out;
  }
  return 0;
}

}  // namespace

extern "C" {

const char* ffgpu_last_error(void) { return g_err.c_str(); }

ffgpu_ctx* ffgpu_create(const int* device_ids, int n_devices) {
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0) {
    g_err = std::string("no CUDA device: ") + cudaGetErrorString(e) +
            " (libffgpu has no CPU evaluation path; a B200 is required)";
    return nullptr;
  }
  std::unique_ptr<ffgpu_ctx> ctx(new ffgpu_ctx);
  std::vector<int> ids;
  if (!device_ids || n_devices <= 0) {
    int cur = 0;
    if (cudaGetDevice(&cur) != cudaSuccess) cur = 0;
    ids.push_back(cur);
  } else {
    ids.assign(device_ids, device_ids + n_devices);
  }
  for (int id : ids) {
    if (id < 0 || id >= count) {
      g_err = "device id " + std::to_string(id) + " out of range";
      return nullptr;
    }
    Device d;
    d.id = id;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, id) == cudaSuccess) d.sm_count = prop.multiProcessorCount;
    // the scratch rows of collections with shared sub-expressions are stream-ordered allocations: keep the pool's
    // memory instead of returning it to the driver at every synchronisation
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, id) == cudaSuccess) {
      uint64_t keep = ~0ULL;
      cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    }
    ctx->devs.push_back(d);
  }
  for (auto& s : ctx->slots) s.dev.resize(ctx->devs.size());
  return ctx.release();
}

void ffgpu_destroy(ffgpu_ctx* ctx) {
  if (!ctx) return;
  for (auto& s : ctx->slots) {
    for (size_t g = 0; g < s.dev.size(); ++g) {
      DeviceGuard guard(ctx->devs[g].id);
      for (auto& b : s.dev[g].buf) {
        if (b.stream) cudaStreamSynchronize(b.stream);
        if (b.d_in) cudaFree(b.d_in);
        if (b.d_out) cudaFree(b.d_out);
        if (b.d_as) cudaFree(b.d_as);
        if (b.h_as) cudaFreeHost(b.h_as);
        if (b.h_stage) cudaFreeHost(b.h_stage);
        if (b.stream) cudaStreamDestroy(b.stream);
      }
    }
  }
  delete ctx;
}

int ffgpu_n_devices(const ffgpu_ctx* ctx) { return ctx ? (int)ctx->devs.size() : 0; }

ffgpu_tape* ffgpu_compile(ffgpu_ctx* ctx, const char* const* functions, size_t n_functions, const char* const* vars,
                          size_t n_vars, unsigned flags) {
  std::unique_ptr<ffgpu_tape> t(new ffgpu_tape);
  std::string err;
  t->col.enable_cse = !(flags & FFGPU_NO_CSE);
  t->col.enable_fuse = !(flags & FFGPU_NO_FUSE);
  t->col.enable_cross = !(flags & FFGPU_NO_CROSS);
  t->col.enable_table = !(flags & FFGPU_NO_TABLE) && !std::getenv("FFGPU_NO_TABLE");
  t->col.enable_split = !(flags & FFGPU_NO_SPLIT) && !std::getenv("FFGPU_NO_SPLIT");
  t->col.enable_horner = !(flags & FFGPU_NO_HORNER) && !std::getenv("FFGPU_NO_HORNER");
  t->col.temp_budget = (int)((flags >> FFGPU_TEMPS_SHIFT) & 15u);
  const bool ok = t->col.compile_strings(to_vec(functions, n_functions), to_vec(vars, n_vars),
                                         (flags & FFGPU_CHECK_IS_EQUAL) != 0, err);
  return finish_compile(ctx, std::move(t), ok, err);
}

ffgpu_tape* ffgpu_compile_text(ffgpu_ctx* ctx, const char* text, const char* const* vars, size_t n_vars,
                               unsigned flags) {
  std::unique_ptr<ffgpu_tape> t(new ffgpu_tape);
  std::string err;
  t->col.enable_cse = !(flags & FFGPU_NO_CSE);
  t->col.enable_fuse = !(flags & FFGPU_NO_FUSE);
  t->col.enable_cross = !(flags & FFGPU_NO_CROSS);
  t->col.enable_table = !(flags & FFGPU_NO_TABLE) && !std::getenv("FFGPU_NO_TABLE");
  t->col.enable_split = !(flags & FFGPU_NO_SPLIT) && !std::getenv("FFGPU_NO_SPLIT");
  t->col.enable_horner = !(flags & FFGPU_NO_HORNER) && !std::getenv("FFGPU_NO_HORNER");
  t->col.temp_budget = (int)((flags >> FFGPU_TEMPS_SHIFT) & 15u);
  const bool ok = t->col.compile_text(text ? text : "", to_vec(vars, n_vars), (flags & FFGPU_CHECK_IS_EQUAL) != 0, err);
  return finish_compile(ctx, std::move(t), ok, err);
}

ffgpu_tape* ffgpu_compile_file(ffgpu_ctx* ctx, const char* path, const char* const* vars, size_t n_vars,
                               unsigned flags) {
  std::ifstream in(path ? path : "");
  if (!in.good()) {
    g_err = std::string("File '") + (path ? path : "") + "' does not exist!";
    return nullptr;
  }
  std::stringstream ss;
  ss << in.rdbuf();
  return ffgpu_compile_text(ctx, ss.str().c_str(), vars, n_vars, flags);
}

ffgpu_tape* ffgpu_tape_clone(const ffgpu_tape* t) {
  if (!t) {
    g_err = "null tape";
    return nullptr;
  }
  std::unique_ptr<ffgpu_tape> n(new ffgpu_tape);
  n->ctx = t->ctx;
  n->col = t->col;
  n->force_k = t->force_k;
  n->force_threads = t->force_threads;
  n->chunk = t->chunk;
  n->coalesce = t->coalesce;
  n->coalesce_cap = t->coalesce_cap;
  if (t->ctx) {
    n->d_tape.assign(t->ctx->devs.size(), nullptr);
    n->d_chunks.assign(t->ctx->devs.size(), nullptr);
    for (int l = 0; l < 2; ++l) {
      n->d_pre[l].assign(t->ctx->devs.size(), nullptr);
      n->d_pre_chunks[l].assign(t->ctx->devs.size(), nullptr);
    }
  }
  // the copy gets device buffers of its own, for the prime the original is resolved for
  if (t->col.prime && !t->col.stale && ffgpu_set_prime(n.get(), t->col.prime)) return nullptr;
  return n.release();
}

void ffgpu_tape_destroy(ffgpu_tape* t) {
  if (!t) return;
  for (CoBatch& b : t->co->b) free_batch(t->ctx, b);
  for (size_t g = 0; g < t->d_tape.size(); ++g) {
    if (!t->d_tape[g]) continue;
    DeviceGuard guard(t->ctx->devs[g].id);
    cudaDeviceSynchronize();
    cudaFree(t->d_tape[g]);
    if (t->d_chunks[g]) cudaFree(t->d_chunks[g]);
    for (int l = 0; l < 2; ++l) {
      if (t->d_pre[l][g]) cudaFree(t->d_pre[l][g]);
      if (t->d_pre_chunks[l][g]) cudaFree(t->d_pre_chunks[l][g]);
    }
  }
  delete t;
}

size_t ffgpu_n_vars(const ffgpu_tape* t) { return t->col.vars.size(); }
size_t ffgpu_n_functions(const ffgpu_tape* t) { return t->col.functions.size(); }
size_t ffgpu_n_out(const ffgpu_tape* t) { return t->col.n_out; }
size_t ffgpu_n_instr(const ffgpu_tape* t) { return t->col.tape.size() - 2; }
size_t ffgpu_stack_depth(const ffgpu_tape* t) { return t->col.stack_depth; }
size_t ffgpu_n_temps(const ffgpu_tape* t) { return t->col.n_temps; }
size_t ffgpu_n_slots(const ffgpu_tape* t) { return t->col.n_slots; }
size_t ffgpu_cse_fallbacks(const ffgpu_tape* t) { return t->col.cse_fallbacks; }
size_t ffgpu_n_structural_duplicates(const ffgpu_tape* t) { return t->col.n_structural_duplicates; }
size_t ffgpu_n_shared_denominators(const ffgpu_tape* t) { return t->col.n_shared_denominators; }

size_t ffgpu_n_table_rows(const ffgpu_tape* t) { return t->col.n_table_rows; }
size_t ffgpu_n_table_levels(const ffgpu_tape* t) { return t->col.pre.size(); }
size_t ffgpu_n_table_rows_level(const ffgpu_tape* t, size_t level) { return level < t->col.pre.size() ? t->col.pre[level].n_out : 0; }
size_t ffgpu_n_split_rows(const ffgpu_tape* t) { return t->col.n_split_rows; }
size_t ffgpu_table_row_offset(const ffgpu_tape* t, size_t level) { return level < t->col.pre.size() ? t->col.pre_row_offset()[level] : 0; }
size_t ffgpu_get_table_tape(const ffgpu_tape* t, size_t level, void* buf, size_t cap) {
  if (level >= t->col.pre.size()) return 0;
  const std::vector<ffgpu::Instr>& tp = t->col.pre[level].tape;
  const size_t n = std::min(cap, tp.size());
  if (buf && n) std::memcpy(buf, tp.data(), n * sizeof(ffgpu::Instr));
  return tp.size();
}

size_t ffgpu_rpn(const ffgpu_tape* t, size_t i, char* buf, size_t cap) {
  if (i >= t->col.functions.size()) return 0;
  std::string s;
  for (const std::string& tok : t->col.functions[i].rpn) {
    if (!s.empty()) s += ' ';
    s += tok;
  }
  if (buf && cap) {
    const size_t n = std::min(s.size(), cap - 1);
    std::memcpy(buf, s.data(), n);
    buf[n] = 0;
  }
  return s.size();
}

size_t ffgpu_get_tape(const ffgpu_tape* t, void* buf, size_t cap) {
  const size_t n = std::min(cap, t->col.tape.size());
  if (buf && n) std::memcpy(buf, t->col.tape.data(), n * sizeof(ffgpu::Instr));
  return t->col.tape.size();
}

void ffgpu_work(const ffgpu_tape* t, uint64_t w[5]) {
  w[0] = t->col.work.tokens;
  w[1] = t->col.work.mm;
  w[2] = t->col.work.ma;
  w[3] = t->col.work.inv;
  w[4] = t->col.work.max_rpn_depth;
}

void ffgpu_tape_work(const ffgpu_tape* t, uint64_t w[5]) {
  w[0] = t->col.executed.tokens;
  w[1] = t->col.executed.mm;
  w[2] = t->col.executed.ma;
  w[3] = t->col.executed.inv;
  w[4] = t->col.executed.max_rpn_depth;
}

uint32_t ffgpu_n_primes(void) { return (uint32_t)prime_table().size(); }
uint64_t ffgpu_prime(uint32_t i) { return i < prime_table().size() ? prime_table()[i] : 0; }
uint64_t ffgpu_get_prime(const ffgpu_tape* t) { return t->col.prime; }

int ffgpu_set_prime(ffgpu_tape* t, uint64_t p) {
  if (!t) return fail("null tape");
  // any odd prime below 2^63 (FFInt::set_new_prime, reference FFInt.cpp:344-353): p = 2^63 - c with c < 2^15 -- all of
  // FireFly's table -- runs on the tuned arithmetic, everything else on the generic (Barrett) kernels.  Primality is
  // the caller's business as it is in the reference, but an even or tiny modulus is certainly a mistake.
  if (p >= (1ULL << 63) || p < 3 || !(p & 1))
    return fail("unsupported modulus " + std::to_string(p) + ": need an odd prime below 2^63");
  std::string err;
  const uint64_t rebuilds = t->col.rebuilds;      // functions added or trimmed, or a tiny prime: the tape has a new size
  if (!t->col.resolve(p, err)) return fail(err);
  if (t->col.rebuilds != rebuilds) drop_device_copies(t);
  t->needs_generic = false;
  for (const ffgpu::Instr& in : t->col.tape) t->needs_generic |= in.fast == ffgpu::FX_GENERIC && in.op != ffgpu::OP_HALT;
  for (size_t l = 0; l < t->col.pre.size() && l < 2; ++l) {
    t->pre_needs_generic[l] = false;
    for (const ffgpu::Instr& in : t->col.pre[l].tape) t->pre_needs_generic[l] |= in.fast == ffgpu::FX_GENERIC && in.op != ffgpu::OP_HALT;
  }
  const size_t bytes = t->col.tape.size() * sizeof(ffgpu::Instr);
  for (size_t g = 0; t->ctx && g < t->ctx->devs.size(); ++g) {
    DeviceGuard guard(t->ctx->devs[g].id);
    // quiescence is the caller's contract; make it true for our own streams anyway
    FF_CUDA(cudaDeviceSynchronize());
    if (!t->d_tape[g]) FF_CUDA(cudaMalloc(&t->d_tape[g], bytes));
    if (!t->d_chunks[g]) {
      FF_CUDA(cudaMalloc(&t->d_chunks[g], t->col.chunk_table.size() * sizeof(uint32_t)));
      FF_CUDA(cudaMemcpy(t->d_chunks[g], t->col.chunk_table.data(), t->col.chunk_table.size() * sizeof(uint32_t),
                         cudaMemcpyHostToDevice));
    }
    FF_CUDA(cudaMemcpy(t->d_tape[g], t->col.tape.data(), bytes, cudaMemcpyHostToDevice));
    for (size_t l = 0; l < t->col.pre.size() && l < 2; ++l) {
      const ffgpu::Collection& pc = t->col.pre[l];
      const size_t pbytes = pc.tape.size() * sizeof(ffgpu::Instr);
      if (!t->d_pre[l][g]) FF_CUDA(cudaMalloc(&t->d_pre[l][g], pbytes));
      if (!t->d_pre_chunks[l][g]) {
        FF_CUDA(cudaMalloc(&t->d_pre_chunks[l][g], pc.chunk_table.size() * sizeof(uint32_t)));
        FF_CUDA(cudaMemcpy(t->d_pre_chunks[l][g], pc.chunk_table.data(), pc.chunk_table.size() * sizeof(uint32_t),
                           cudaMemcpyHostToDevice));
      }
      FF_CUDA(cudaMemcpy(t->d_pre[l][g], pc.tape.data(), pbytes, cudaMemcpyHostToDevice));
    }
    // cudaMemcpy from pageable memory returns when the data sits in the driver's staging buffer, not when it has
    // reached the device, and the evaluation streams are non-blocking (no implicit order with the copies' stream):
    // without this wait a launch right after ffgpu_set_prime could read a tape that is still arriving (one wrong
    // value or illegal access in about 5000 compile-and-evaluate rounds of scripts/fuzz_gpu.py)
    FF_CUDA(cudaDeviceSynchronize());
  }
  return 0;
}

int ffgpu_set_factors(ffgpu_tape* t, const uint32_t* out_index, const char* const* factors, size_t n_factors,
                      const char* const* factor_vars, size_t n_factor_vars) {
  if (!t) return fail("null tape");
  std::vector<std::pair<uint32_t, std::string>> fac;
  for (size_t i = 0; i < n_factors; ++i) fac.emplace_back(out_index[i], factors[i] ? factors[i] : "");
  std::string err;
  if (!t->col.apply_factors(fac, to_vec(factor_vars, n_factor_vars), err)) return fail(err);
  // the tape changed size: drop the device copies, ffgpu_set_prime uploads the new one
  drop_device_copies(t);
  return 0;
}

long ffgpu_add_function(ffgpu_tape* t, const char* function, unsigned flags) {
  if (!t) return fail("null tape");
  uint64_t pos = 0;
  std::string err;
  const bool plain = (flags & FFGPU_ADD_PLAIN) != 0;
  if (!t->col.append(function ? function : "", plain ? (flags & FFGPU_ADD_VALIDATE) != 0 : true, !plain,
                     (flags & FFGPU_ADD_NO_DUPLICATES) != 0, pos, err))
    return fail(err);
  return (long)pos;
}

long ffgpu_add_function_rpn(ffgpu_tape* t, const char* const* tokens, size_t n_tokens) {
  if (!t) return fail("null tape");
  uint64_t pos = 0;
  std::string err;
  if (!t->col.append_rpn(to_vec(tokens, n_tokens), pos, err)) return fail(err);
  return (long)pos;
}

int ffgpu_trim(ffgpu_tape* t, const uint64_t* keep, size_t n_keep, uint64_t* new_positions) {
  if (!t) return fail("null tape");
  std::vector<uint64_t> k(keep, keep + n_keep), np;
  std::string err;
  if (!t->col.trim(k, np, err)) return fail(err);
  if (new_positions) std::copy(np.begin(), np.end(), new_positions);
  return 0;
}

int ffgpu_is_stale(const ffgpu_tape* t) { return t && t->col.stale ? 1 : 0; }

int ffgpu_set_launch(ffgpu_tape* t, int k, int threads) {
  if (!(k == 0 || k == 1 || k == 2 || k == 4 || k == 8)) return fail("probes_per_thread must be 0, 1, 2, 4 or 8");
  if (!(threads == 0 || threads == 128 || threads == 256 || threads == 512 || threads == 1024))
    return fail("threads_per_block must be 0, 128, 256, 512 or 1024");
  t->force_k = k;
  t->force_threads = threads;
  return 0;
}

int ffgpu_set_coalesce(ffgpu_tape* t, int enable, size_t max_points) {
  if (!t) return fail("null tape");
  if (max_points > (1u << 20)) return fail("at most 2^20 points per coalesced batch");
  t->coalesce = enable < 0 ? -1 : (enable ? 1 : 0);
  t->coalesce_cap = max_points;
  return 0;
}

void ffgpu_coalesce_stats(const ffgpu_tape* t, uint64_t s[3]) {
  std::lock_guard<std::mutex> lk(t->co->m);
  s[0] = t->co->batches;
  s[1] = t->co->calls;
  s[2] = t->co->points;
}

int ffgpu_set_chunk(ffgpu_tape* t, size_t points) {
  t->chunk = points;
  return 0;
}

uint64_t ffgpu_launch_count(void) { return g_launches.load(); }

int ffgpu_eval_device(ffgpu_tape* t, int g, const uint64_t* d_values, size_t in_stride, size_t n_points,
                      uint64_t* d_out, size_t out_stride, void* stream) {
  if (!t) return fail("null tape");
  if (!t->ctx) return fail("tape was compiled without a context: no GPU to evaluate on (there is no CPU path)");
  if (g < 0 || g >= (int)t->ctx->devs.size()) return fail("device index out of range");
  if (t->col.stale) return fail("functions were added or trimmed: call ffgpu_set_prime before evaluating");
  if (!t->col.prime || !t->d_tape[g]) return fail("ffgpu_set_prime has not been called");
  if (in_stride < n_points || out_stride < n_points) return fail("row stride smaller than n_points");
  if (n_points == 0) return 0;
  DeviceGuard guard(t->ctx->devs[g].id);
  return launch(t, g, d_values, in_stride, n_points, d_out, out_stride, (cudaStream_t)stream);
}

}  // extern "C"

namespace {
inline void cpu_relax() {
#if defined(__x86_64__) || defined(__i386__)
  __builtin_ia32_pause();
#else
  std::this_thread::yield();
#endif
}

int coalesce_default() {
  static const int v = [] {
    const char* e = std::getenv("FFGPU_COALESCE");
    // Off unless asked for: on a B200 box the unmodified Reconstructor sees a LOWER black-box time per probe when every
    // call launches on its own slot's stream (32 slots) than when concurrent calls wait for each other to share a
    // launch -- f4, 16 threads: 2.6e-5 s against 3.0e-5 s; s_y_4_v, 4 threads: 4.7e-4 against 7.6e-4
    // (profiles/r02_reconstruct_aggregator_on_off.log).  What the aggregator buys is an eighth to a twelfth of the launches.
    return e && *e ? std::atoi(e) : 0;
  }();
  return v;
}

// (re)allocate the staging of one batch for `cap` points of the tape's current shape on device `dev`
int prepare_batch(ffgpu_tape* t, CoBatch& b, size_t cap, int dev) {
  const size_t n_vars = t->col.vars.size(), n_out = t->col.n_out;
  if (b.cap == cap && b.n_vars == n_vars && b.n_out == n_out && b.dev == dev) return 0;
  free_batch(t->ctx, b);
  DeviceGuard guard(t->ctx->devs[dev].id);
  const size_t w_in = std::max<size_t>(n_vars * cap, 1), w_out = std::max<size_t>(n_out * cap, 1);
  FF_CUDA(cudaHostAlloc(&b.h_in, (w_in + w_out) * sizeof(uint64_t), cudaHostAllocPortable | cudaHostAllocMapped));
  b.h_out = b.h_in + w_in;
  FF_CUDA(cudaMalloc(&b.d_in, w_in * sizeof(uint64_t)));
  FF_CUDA(cudaStreamCreateWithFlags(&b.stream, cudaStreamNonBlocking));
  b.cap = cap;
  b.n_vars = n_vars;
  b.n_out = n_out;
  b.dev = dev;
  return 0;
}

// One small call through the aggregator (struct CoBatch).  Returns like ffgpu_eval.
int coalesced_eval(ffgpu_tape* t, const uint64_t* values, size_t n_points, uint64_t* out) {
  Coalescer& co = *t->co;
  const size_t n_vars = t->col.vars.size(), n_out = t->col.n_out;
  const size_t G = t->ctx->devs.size();
  size_t cap = t->coalesce_cap ? t->coalesce_cap : std::min<size_t>(4096, std::max<size_t>(512, (size_t)(128 * 1024) / (n_vars + n_out + 1)));
  cap = (std::max(cap, n_points) + 1) & ~(size_t)1;
  std::unique_lock<std::mutex> lk(co.m);
  CoBatch* B = nullptr;
  for (;;) {
    B = &co.b[co.open];
    if (!B->closed && B->users == 0 && B->count == 0) {
      // first caller of a fresh batch: make sure its buffers fit the tape's shape (it may have been rebuilt)
      if (prepare_batch(t, *B, std::max(cap, B->cap), (int)((size_t)co.open % G))) return -1;
      B->seq = co.next_seq++;
      B->done = false;
      B->has_leader = false;
      B->rc = 0;
    }
    if (!B->closed && B->count + n_points <= B->cap) break;
    co.cv.wait(lk);                       // the open batch is full or already on its way: wait for the next one
  }
  const size_t off = B->count;
  B->count += n_points;
  ++B->users;
  ++co.calls;
  co.points += n_points;
  for (size_t v = 0; v < n_vars; ++v) std::memcpy(B->h_in + v * B->cap + off, values + v * n_points, n_points * sizeof(uint64_t));
  const uint64_t my_seq = B->seq;
  if (!B->has_leader) {
    B->has_leader = true;
    // leader: wait for the GPU (callers arriving meanwhile join this batch), then close and run it
    co.cv.wait(lk, [&] { return !co.busy && co.b[co.open ^ 1].users == 0; });
    B->closed = true;
    co.busy = true;
    co.open ^= 1;
    CoBatch& nb = co.b[co.open];          // the other batch has been taken by all its callers: it fills next
    nb.count = 0;
    nb.closed = false;
    nb.done = false;
    nb.has_leader = false;
    ++co.batches;
    const size_t cnt = B->count, cnt_even = (cnt + 1) & ~(size_t)1;
    lk.unlock();
    co.cv.notify_all();
    int prev = 0;
    cudaGetDevice(&prev);
    const int rc = [&]() -> int {
      FF_CUDA(cudaSetDevice(t->ctx->devs[B->dev].id));
      if (cnt_even != cnt)
        for (size_t v = 0; v < n_vars; ++v) B->h_in[v * B->cap + cnt] = 0;       // the pad lane holds a residue
      if (n_vars)
        FF_CUDA(cudaMemcpy2DAsync(B->d_in, B->cap * 8, B->h_in, B->cap * 8, cnt_even * 8, n_vars, cudaMemcpyHostToDevice, B->stream));
      // results are written by the kernel straight into the mapped pinned staging (no copy back)
      if (launch(t, B->dev, B->d_in, B->cap, cnt_even, B->h_out, B->cap, B->stream)) return -1;
      FF_CUDA(cudaStreamSynchronize(B->stream));
      return 0;
    }();
    cudaSetDevice(prev);
    lk.lock();
    B->rc = rc;
    if (rc) B->err = g_err;
    B->done = true;
    B->done_seq.store(my_seq, std::memory_order_release);
    co.busy = false;
    co.cv.notify_all();
  } else {
    // a launch takes tens of microseconds, a sleeping thread longer than that to wake up: look for the result a few
    // thousand times before going to sleep on the condition variable
    lk.unlock();
    for (int spin = 0; spin < 20000 && B->done_seq.load(std::memory_order_acquire) != my_seq; ++spin) cpu_relax();
    lk.lock();
    co.cv.wait(lk, [&] { return B->done && B->seq == my_seq; });
  }
  const int rc = B->rc;
  if (rc) g_err = B->err;
  lk.unlock();
  // the batch is not reused before every caller has left it (users), so the slice is read without the lock
  if (rc == 0)
    for (size_t f = 0; f < n_out; ++f) std::memcpy(out + f * n_points, B->h_out + f * B->cap + off, n_points * sizeof(uint64_t));
  lk.lock();
  --B->users;
  lk.unlock();
  co.cv.notify_all();
  return rc;
}

// Host-buffer evaluation shared by ffgpu_eval (values given) and ffgpu_eval_line (values == NULL: the
// points are built on the device from t[], a[], s[]).
int eval_host(ffgpu_tape* t, const uint64_t* values, const uint64_t* line_a, const uint64_t* line_s,
              const uint64_t* line_t, size_t n_points, uint64_t* out, uint32_t thread_id) {
  if (!t) return fail("null tape");
  if (!t->ctx) return fail("tape was compiled without a context: no GPU to evaluate on (there is no CPU path)");
  if (t->col.stale) return fail("functions were added or trimmed: call ffgpu_set_prime before evaluating");
  if (!t->col.prime) return fail("ffgpu_set_prime has not been called");
  if (n_points == 0) return 0;
  ffgpu_ctx* ctx = t->ctx;
  const size_t G = ctx->devs.size();
  const size_t n_vars = t->col.vars.size(), n_out = t->col.n_out;
  // Small calls -- the Reconstructor's bunches of 1..128 points (Reconstructor.hpp:3330-3354) -- are latency
  // bound; concurrent ones share a launch (coalesced_eval)
  constexpr size_t SMALL_WORDS = 32 * 1024;            // 256 KB in + out
  const bool small = values && (n_vars + n_out) * (n_points + 1) <= SMALL_WORDS;
  if (small && (t->coalesce >= 0 ? t->coalesce : coalesce_default()) && n_vars + n_out > 0)
    return coalesced_eval(t, values, n_points, out);
  Slot& slot = ctx->slots[thread_id % N_SLOTS];
  std::lock_guard<std::mutex> lock(slot.mtx);

  // without the aggregator: one GPU, one contiguous pinned copy each way instead of strided copies from pageable
  // memory, one launch per call
  if (small) {
    int prev = 0;
    cudaGetDevice(&prev);
    const int rc_small = [&]() -> int {
      FF_CUDA(cudaSetDevice(ctx->devs[0].id));
      DevBuf& b = slot.dev[0].buf[0];
      const size_t n_even = (n_points + 1) & ~(size_t)1;
      const size_t w_in = std::max<size_t>(n_vars * n_even, 1), w_out = std::max<size_t>(n_out * n_even, 1);
      if (ensure(b, w_in + n_even, w_out)) return -1;
      if (b.cap_stage < w_in + w_out) {
        if (b.h_stage) FF_CUDA(cudaFreeHost(b.h_stage));
        b.h_stage = nullptr;
        b.cap_stage = 0;
        FF_CUDA(cudaHostAlloc(&b.h_stage, std::max<size_t>(w_in + w_out, SMALL_WORDS + 64) * sizeof(uint64_t), cudaHostAllocPortable | cudaHostAllocMapped));
        b.cap_stage = std::max<size_t>(w_in + w_out, SMALL_WORDS + 64);
      }
      uint64_t* h_in = b.h_stage;
      uint64_t* h_out = b.h_stage + w_in;
      for (size_t v = 0; v < n_vars; ++v) {
        std::memcpy(h_in + v * n_even, values + v * n_points, n_points * sizeof(uint64_t));
        if (n_even != n_points) h_in[v * n_even + n_points] = 0;
      }
      // The staging buffer is pinned, mapped host memory (one address space with the device): the kernel can
      // read the coordinates from it and write the results into it directly over PCIe, which saves a copy
      // engine round trip each way.  zero_copy: 0 = copy both ways, 1 = results written by the kernel,
      // 2 = coordinates read by the kernel as well.
      const int zc = small_zero_copy();
      const uint64_t* k_in = b.d_in;
      uint64_t* k_out = b.d_out;
      if (zc >= 2) k_in = h_in;
      else if (n_vars) FF_CUDA(cudaMemcpyAsync(b.d_in, h_in, n_vars * n_even * sizeof(uint64_t), cudaMemcpyHostToDevice, b.stream));
      if (zc >= 1) k_out = h_out;
      if (launch(t, 0, k_in, n_even, n_even, k_out, n_even, b.stream)) return -1;
      if (zc < 1 && n_out) FF_CUDA(cudaMemcpyAsync(h_out, b.d_out, n_out * n_even * sizeof(uint64_t), cudaMemcpyDeviceToHost, b.stream));
      FF_CUDA(cudaStreamSynchronize(b.stream));
      for (size_t f = 0; f < n_out; ++f) std::memcpy(out + f * n_points, h_out + f * n_even, n_points * sizeof(uint64_t));
      return 0;
    }();
    cudaSetDevice(prev);
    return rc_small;
  }

  // contiguous shard per device; chunk size: large enough to amortise launches, small enough to pipeline
  const size_t per_dev = (n_points + G - 1) / G;
  size_t chunk = t->chunk;
  if (!chunk) {
    // bytes that cross PCIe per point: the line form ships t only
    const size_t bytes_per_point = 8 * ((values ? n_vars : 1) + n_out);
    chunk = std::max<size_t>(1 << 16, ((values ? 32u : 6u) << 20) / std::max<size_t>(bytes_per_point, 1));
  }
  chunk = std::min(chunk, per_dev);
  chunk = (chunk + 1) & ~(size_t)1;
  const size_t n_chunks = (per_dev + chunk - 1) / chunk;
  int prev_dev = 0;
  cudaGetDevice(&prev_dev);
  int rc = 0;
  for (size_t ci = 0; ci < n_chunks && rc == 0; ++ci) {
    for (size_t g = 0; g < G && rc == 0; ++g) {
      const size_t lo_dev = g * per_dev, hi_dev = std::min(n_points, lo_dev + per_dev);
      const size_t lo = lo_dev + ci * chunk;
      if (lo >= hi_dev) continue;
      const size_t cnt = std::min(chunk, hi_dev - lo);
      const size_t cnt_even = (cnt + 1) & ~(size_t)1;     // the pad lane computes on junk and is never copied back
      rc = [&]() -> int {
        FF_CUDA(cudaSetDevice(ctx->devs[g].id));
        DevBuf& b = slot.dev[g].buf[ci % N_BUF];
        if (ensure(b, chunk * (n_vars + 1), chunk * std::max<size_t>(n_out, 1))) return -1;
        if (values) {
          if (n_vars) {
            if (copy_rows(b.d_in, chunk, values + lo, n_points, cnt, n_vars, cudaMemcpyHostToDevice, b.stream)) return -1;
            // the pad lane of an odd chunk must hold a residue: its divisor enters the block's product tree
            if (cnt_even != cnt) FF_CUDA(cudaMemset2DAsync(b.d_in + cnt, chunk * 8, 0, 8, n_vars, b.stream));
          }
        } else if (n_vars) {
          // t travels, the coordinates are built here: row n_vars of the staging buffer holds t
          if (b.cap_as < 2 * n_vars) {
            FF_CUDA(cudaStreamSynchronize(b.stream));
            if (b.d_as) FF_CUDA(cudaFree(b.d_as));
            if (b.h_as) FF_CUDA(cudaFreeHost(b.h_as));
            b.d_as = nullptr;
            b.h_as = nullptr;
            b.cap_as = 0;
            FF_CUDA(cudaMalloc(&b.d_as, 2 * n_vars * sizeof(uint64_t)));
            FF_CUDA(cudaHostAlloc(&b.h_as, 2 * n_vars * sizeof(uint64_t), cudaHostAllocPortable));
            b.cap_as = 2 * n_vars;
          }
          uint64_t* d_t = b.d_in + chunk * n_vars;
          if (ci < (size_t)N_BUF) {
            // first use of this buffer in this call: a and s go up once, from pinned memory (a copy from the
            // caller's pageable arrays would synchronise the stream at every chunk and stall the pipeline)
            std::memcpy(b.h_as, line_a, n_vars * 8);
            std::memcpy(b.h_as + n_vars, line_s, n_vars * 8);
            FF_CUDA(cudaMemcpyAsync(b.d_as, b.h_as, 2 * n_vars * 8, cudaMemcpyHostToDevice, b.stream));
          }
          FF_CUDA(cudaMemcpyAsync(d_t, line_t + lo, cnt * 8, cudaMemcpyHostToDevice, b.stream));
          if (cnt_even != cnt) FF_CUDA(cudaMemsetAsync(d_t + cnt, 0, 8, b.stream));
          cudaError_t e = ffgpu::launch_line_points(b.d_in, d_t, b.d_as, b.d_as + n_vars, (uint32_t)n_vars, cnt_even,
                                                    chunk, t->col.prime, b.stream);
          if (e != cudaSuccess) return cuda_fail(e, "line_points");
          g_launches.fetch_add(1, std::memory_order_relaxed);
        }
        if (launch(t, (int)g, b.d_in, chunk, cnt_even, b.d_out, chunk, b.stream)) return -1;
        if (n_out && copy_rows(out + lo, n_points, b.d_out, chunk, cnt, n_out, cudaMemcpyDeviceToHost, b.stream)) return -1;
        return 0;
      }();
    }
  }
  // drain
  for (size_t g = 0; g < G; ++g) {
    cudaSetDevice(ctx->devs[g].id);
    for (auto& b : slot.dev[g].buf) {
      if (!b.stream) continue;
      cudaError_t e = cudaStreamSynchronize(b.stream);
      if (e != cudaSuccess && rc == 0) rc = cuda_fail(e, "cudaStreamSynchronize");
    }
  }
  cudaSetDevice(prev_dev);
  return rc;
}
}  // namespace

extern "C" {

int ffgpu_eval(ffgpu_tape* t, const uint64_t* values, size_t n_points, uint64_t* out, uint32_t thread_id) {
  if (t && n_points && t->col.vars.size() && !values) return fail("null values");
  static const uint64_t none = 0;
  return eval_host(t, values ? values : &none, nullptr, nullptr, nullptr, n_points, out, thread_id);
}

int ffgpu_eval_line(ffgpu_tape* t, const uint64_t* a, const uint64_t* s, const uint64_t* tt, size_t n_points,
                    uint64_t* out, uint32_t thread_id) {
  if (t && n_points && (!tt || (t->col.vars.size() && (!a || !s)))) return fail("null line parameters");
  if (t && t->col.prime && n_points) {
    for (size_t v = 0; v < t->col.vars.size(); ++v)
      if (a[v] >= t->col.prime || s[v] >= t->col.prime) return fail("line parameters must be residues below the prime");
  }
  return eval_host(t, nullptr, a, s, tt, n_points, out, thread_id);
}

void* ffgpu_host_alloc(size_t bytes) {
  void* p = nullptr;
  cudaError_t e = cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocPortable);
  if (e != cudaSuccess) {
    cuda_fail(e, "cudaHostAlloc");
    return nullptr;
  }
  return p;
}
void ffgpu_host_free(void* p) {
  if (p) cudaFreeHost(p);
}

int ffgpu_fill_points(ffgpu_ctx* ctx, int g, uint64_t* d_values, size_t n_vars, size_t n_points, size_t stride,
                      uint64_t first_point, uint64_t total_points, uint64_t seed, uint64_t p, void* stream) {
  if (!ctx || g < 0 || g >= (int)ctx->devs.size()) return fail("device index out of range");
  DeviceGuard guard(ctx->devs[g].id);
  cudaError_t e = ffgpu::launch_fill_points(d_values, n_vars, n_points, stride, first_point, total_points, seed, p,
                                            (cudaStream_t)stream);
  if (e != cudaSuccess) return cuda_fail(e, "fill_points");
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return 0;
}

int ffgpu_checksum(ffgpu_ctx* ctx, int g, const uint64_t* d_out, size_t n_out, size_t n_points, size_t stride,
                   uint64_t p, uint64_t* d_sums, void* stream) {
  if (!ctx || g < 0 || g >= (int)ctx->devs.size()) return fail("device index out of range");
  DeviceGuard guard(ctx->devs[g].id);
  cudaError_t e = ffgpu::launch_checksum(d_out, n_out, n_points, stride, p, d_sums, (cudaStream_t)stream);
  if (e != cudaSuccess) return cuda_fail(e, "checksum");
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return 0;
}

int ffgpu_field_ops(ffgpu_ctx* ctx, int op, const uint64_t* a, const uint64_t* b, uint64_t* r, size_t n, uint64_t p) {
  if (!ctx) return fail("null context");
  if (n == 0) return 0;
  if (p >= (1ULL << 63) || p < 3 || !(p & 1)) return fail("unsupported modulus: need an odd prime below 2^63");
  DeviceGuard guard(ctx->devs[0].id);
  uint64_t *da = nullptr, *db = nullptr, *dr = nullptr;
  cudaError_t e = cudaMalloc(&da, n * 8);
  if (e == cudaSuccess) e = cudaMalloc(&db, n * 8);
  if (e == cudaSuccess) e = cudaMalloc(&dr, n * 8);
  if (e == cudaSuccess) e = cudaMemcpy(da, a, n * 8, cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(db, b, n * 8, cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = ffgpu::launch_field_ops(op, da, db, dr, n, p, nullptr);
  if (e == cudaSuccess) e = cudaDeviceSynchronize();
  if (e == cudaSuccess) e = cudaMemcpy(r, dr, n * 8, cudaMemcpyDeviceToHost);
  cudaFree(da);
  cudaFree(db);
  cudaFree(dr);
  if (e != cudaSuccess) return cuda_fail(e, "field_ops");
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return 0;
}

}  // extern "C"
