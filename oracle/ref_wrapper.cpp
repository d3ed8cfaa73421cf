// ref_wrapper.cpp -- TEST INFRASTRUCTURE ONLY.
//
// Thin extern "C" facade over the UNMODIFIED reference (jklappert/FireFly 2.0.3,
// compiled from /root/reference by oracle/Makefile into oracle/_ref/) so that
// tests and the CPU baseline of bench.py can drive the reference's own
// ShuntingYardParser::evaluate_pre (reference:
// source/include/firefly/ShuntingYardParser.hpp:365-479) from ctypes.
// Nothing in firefly_b200/ links or loads this.
//
// FFInt::p is a process-wide static (reference: source/FFInt.cpp:25): one prime
// at a time; ffref_set_prime must not race with evaluations.
#include <cstdint>
#include <cstring>
#include <string>
#include <thread>
#include <unordered_map>
#include <unordered_set>
#include <vector>

#include "firefly/FFInt.hpp"
#include "firefly/FFIntVec.hpp"
#include "firefly/ReconstHelper.hpp"
#include "firefly/ShuntingYardParser.hpp"
#include "firefly/utils.hpp"

using namespace firefly;

struct ffref_parser {
  ShuntingYardParser par;
  size_t n_vars;
  size_t n_out;
};

static std::vector<std::string> to_vec(const char* const* s, size_t n) {
  std::vector<std::string> v;
  v.reserve(n);
  for (size_t i = 0; i < n; ++i) v.emplace_back(s[i]);
  return v;
}

template <int N>
static void eval_block(const ShuntingYardParser& par, size_t n_vars, size_t n_out, const uint64_t* values,
                       size_t stride_in, uint64_t* out, size_t stride_out, size_t j0) {
  std::vector<FFIntVec<N>> in(n_vars);
  for (size_t v = 0; v < n_vars; ++v)
    for (int l = 0; l < N; ++l) in[v][l].n = values[v * stride_in + j0 + l];
  std::vector<FFIntVec<N>> res = par.evaluate_pre(in);
  for (size_t f = 0; f < n_out; ++f)
    for (int l = 0; l < N; ++l) out[f * stride_out + j0 + l] = res[f][l].n;
}

extern "C" {

void ffref_set_prime(uint64_t p) { FFInt::set_new_prime(p); }
uint64_t ffref_get_prime() { return FFInt::p; }
uint64_t ffref_prime(uint32_t i) { return primes()[i]; }
uint32_t ffref_n_primes() { return (uint32_t)primes().size(); }

// functions given as strings (reference ctor: ShuntingYardParser.cpp:45-58)
ffref_parser* ffref_parse(const char* const* funs, size_t n_fun, const char* const* vars, size_t n_vars,
                          int check_is_equal) {
  ffref_parser* h = new ffref_parser{ShuntingYardParser(to_vec(funs, n_fun), to_vec(vars, n_vars), check_is_equal != 0, true),
                                     n_vars, 0};
  std::vector<FFInt> probe(n_vars, FFInt(1));
  h->n_out = h->par.evaluate_pre(probe).size();
  return h;
}

// functions read from a ';'-separated file (reference ctor: ShuntingYardParser.cpp:32-43)
ffref_parser* ffref_parse_file(const char* file, const char* const* vars, size_t n_vars, int check_is_equal) {
  ffref_parser* h = new ffref_parser{ShuntingYardParser(std::string(file), to_vec(vars, n_vars), check_is_equal != 0, true),
                                     n_vars, 0};
  std::vector<FFInt> probe(n_vars, FFInt(1));
  h->n_out = h->par.evaluate_pre(probe).size();
  return h;
}

void ffref_free(ffref_parser* h) { delete h; }
size_t ffref_n_out(const ffref_parser* h) { return h->n_out; }
size_t ffref_n_functions(const ffref_parser* h) { return h->par.get_size(); }

// what BlackBox::prime_changed() does (reference: benchmarks/benchmarks.cpp:31-33)
void ffref_prime_changed(ffref_parser* h) { h->par.precompute_tokens(); }
void ffref_precompute_force(ffref_parser* h) { h->par.precompute_tokens(true); }

// --- the rest of the parser's public interface (reference: ShuntingYardParser.cpp:315-377, 720-776) ---
size_t ffref_add_otf(ffref_parser* h, const char* fun, int no_duplicates) {
  std::string f(fun);
  return h->par.add_otf(f, no_duplicates != 0);
}
void ffref_parse_function(ffref_parser* h, const char* fun, const char* const* vars, size_t n_vars, int validate) {
  std::string f(fun);
  h->par.parse_function(f, to_vec(vars, n_vars), validate != 0);
}
size_t ffref_add_otf_precompute(ffref_parser* h, const char* const* tokens, size_t n) {
  return h->par.add_otf_precompute(to_vec(tokens, n));
}
// new_pos[k] = new position of keep[k], ~0 if it was dropped
void ffref_trim(ffref_parser* h, const uint64_t* keep, size_t n, uint64_t* new_pos) {
  std::unordered_set<size_t> k(keep, keep + n);
  std::unordered_map<size_t, size_t> m = h->par.trim(k);
  for (size_t i = 0; i < n; ++i) new_pos[i] = m.count(keep[i]) ? (uint64_t)m[keep[i]] : ~0ULL;
}
// after a change: rebuild the precomputed tokens (force) and count the outputs again
void ffref_refresh(ffref_parser* h, int force) {
  if (force) h->par.precompute_tokens(true);
  std::vector<FFInt> probe(h->n_vars, FFInt(1));
  h->n_out = h->par.evaluate_pre(probe).size();
}

// RPN of function i, tokens joined by ' ' (reference: get_rp_function, ShuntingYardParser.cpp:534-536).
// Returns the length needed (without NUL); writes at most cap-1 chars.
size_t ffref_rpn(const ffref_parser* h, size_t i, char* buf, size_t cap) {
  std::vector<std::string> t = h->par.get_rp_function(i);
  std::string s;
  for (size_t k = 0; k < t.size(); ++k) {
    if (k) s += ' ';
    s += t[k];
  }
  if (cap) {
    size_t n = s.size() < cap - 1 ? s.size() : cap - 1;
    std::memcpy(buf, s.data(), n);
    buf[n] = 0;
  }
  return s.size();
}

// scalar evaluate_pre<FFInt>, one point per call.  values: [n_vars][n_points], out: [n_out][n_points]
void ffref_eval_scalar(const ffref_parser* h, const uint64_t* values, size_t n_points, uint64_t* out) {
  std::vector<FFInt> in(h->n_vars);
  for (size_t j = 0; j < n_points; ++j) {
    for (size_t v = 0; v < h->n_vars; ++v) in[v].n = values[v * n_points + j];
    std::vector<FFInt> res = h->par.evaluate_pre(in);
    for (size_t f = 0; f < h->n_out; ++f) out[f * n_points + j] = res[f].n;
  }
}

// bunched evaluate_pre<FFIntVec<N>> over [j0, j1), falling to smaller bunches at the tail.
static void eval_range(const ffref_parser* h, const uint64_t* values, size_t n_points, uint64_t* out, size_t j0,
                       size_t j1, int bunch) {
  size_t j = j0;
  while (j < j1) {
    size_t left = j1 - j;
    if (bunch >= 128 && left >= 128) { eval_block<128>(h->par, h->n_vars, h->n_out, values, n_points, out, n_points, j); j += 128; }
    else if (bunch >= 64 && left >= 64) { eval_block<64>(h->par, h->n_vars, h->n_out, values, n_points, out, n_points, j); j += 64; }
    else if (bunch >= 32 && left >= 32) { eval_block<32>(h->par, h->n_vars, h->n_out, values, n_points, out, n_points, j); j += 32; }
    else if (bunch >= 16 && left >= 16) { eval_block<16>(h->par, h->n_vars, h->n_out, values, n_points, out, n_points, j); j += 16; }
    else if (bunch >= 8 && left >= 8) { eval_block<8>(h->par, h->n_vars, h->n_out, values, n_points, out, n_points, j); j += 8; }
    else if (bunch >= 4 && left >= 4) { eval_block<4>(h->par, h->n_vars, h->n_out, values, n_points, out, n_points, j); j += 4; }
    else if (bunch >= 2 && left >= 2) { eval_block<2>(h->par, h->n_vars, h->n_out, values, n_points, out, n_points, j); j += 2; }
    else {
      std::vector<FFInt> in(h->n_vars);
      for (size_t v = 0; v < h->n_vars; ++v) in[v].n = values[v * n_points + j];
      std::vector<FFInt> res = h->par.evaluate_pre(in);
      for (size_t f = 0; f < h->n_out; ++f) out[f * n_points + j] = res[f].n;
      j += 1;
    }
  }
}

// the reference's multithreaded CPU path: n_threads workers, each evaluating a
// contiguous slice in bunches of `bunch` (1,2,4,...,128) like Reconstructor's
// thread pool does (reference: Reconstructor.hpp:3325-3477).
void ffref_eval(const ffref_parser* h, const uint64_t* values, size_t n_points, uint64_t* out, int bunch,
                int n_threads) {
  if (n_threads <= 1) {
    eval_range(h, values, n_points, out, 0, n_points, bunch);
    return;
  }
  std::vector<std::thread> th;
  size_t per = (n_points + n_threads - 1) / n_threads;
  per = (per + 127) / 128 * 128;
  for (int t = 0; t < n_threads; ++t) {
    size_t j0 = (size_t)t * per, j1 = j0 + per < n_points ? j0 + per : n_points;
    if (j0 >= j1) break;
    th.emplace_back([=] { eval_range(h, values, n_points, out, j0, j1, bunch); });
  }
  for (auto& t : th) t.join();
}

// bunch policy (reference: source/utils.cpp:305-344)
uint32_t ffref_compute_bunch_size(uint32_t queue_length, uint32_t thr_n, uint32_t max_bunch) {
  return compute_bunch_size(queue_length, thr_n, max_bunch);
}
uint32_t ffref_compute_job_number(uint32_t queue_length, uint32_t threads, uint32_t total_threads, uint32_t max_bunch) {
  return compute_job_number(queue_length, threads, total_threads, max_bunch);
}

// field ops straight from FFInt (DEFAULT arithmetic), for arithmetic parity tests
uint64_t ffref_add(uint64_t a, uint64_t b) { return (FFInt(a) + FFInt(b)).n; }
uint64_t ffref_sub(uint64_t a, uint64_t b) { return (FFInt(a) - FFInt(b)).n; }
uint64_t ffref_mul(uint64_t a, uint64_t b) { return (FFInt(a) * FFInt(b)).n; }
uint64_t ffref_div(uint64_t a, uint64_t b) { return (FFInt(a) / FFInt(b)).n; }
uint64_t ffref_neg(uint64_t a) { return (-FFInt(a)).n; }
uint64_t ffref_pow(uint64_t a, uint64_t e) { return FFInt(a).pow(FFInt(e)).n; }
uint64_t ffref_pow_int(uint64_t a, int e) { return FFInt(a).pow(e).n; }
int ffref_to_neg_int(uint64_t a) { return FFInt(a).to_neg_int(); }

}  // extern "C"
