// pipe_peak.cu -- integer-pipe microbenchmark for B200 (sm_100a).
//
// MEASURED_PEAKS.json carries HBM and bf16 peaks only; the hot path of this
// repo (mod-p probe evaluation) is bounded by the integer multiply pipe, so the
// roofline denominator has to be measured: independent chains of 32-bit
// mad.lo / mad.hi / mad.wide, plus IADD3 / LOP3 / SHF and a few mixes that tell
// how the fma and alu pipes share issue slots.  Output: one JSON object.
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o pipe_peak pipe_peak.cu
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <vector>
#include <string>
#include <cuda_runtime.h>

typedef unsigned long long u64;
typedef unsigned int u32;

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { \
  fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(2); } } while (0)

constexpr int CHAINS = 8;     // independent dependency chains per thread
constexpr int INNER = 16;     // unrolled repeats per loop trip

// Each kernel: CHAINS independent accumulators, INNER*CHAINS instructions of the
// probed kind per trip.  Operands come from memory so nothing constant-folds.
#define PIPE_KERNEL(NAME, DECL, INIT, BODY, FINI)                                   \
  __global__ void __launch_bounds__(256) NAME(u32* out, const u32* in, int trips) { \
    u32 a = in[threadIdx.x & 31], b = in[32 + (threadIdx.x & 31)];                  \
    DECL;                                                                           \
    _Pragma("unroll") for (int k = 0; k < CHAINS; k++) { INIT; }                    \
    _Pragma("unroll 1") for (int t = 0; t < trips; t++) {                           \
      _Pragma("unroll") for (int i = 0; i < INNER; i++) {                           \
        _Pragma("unroll") for (int k = 0; k < CHAINS; k++) { BODY; }                \
      }                                                                             \
    }                                                                               \
    u32 r = 0;                                                                      \
    _Pragma("unroll") for (int k = 0; k < CHAINS; k++) { FINI; }                    \
    if (r == 0x12345678u) out[threadIdx.x] = r;                                     \
  }

// 32-bit IMAD (mad.lo.u32): x = x*a + b
PIPE_KERNEL(k_imad_lo, u32 x[CHAINS], x[k] = a + k,
            asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x[k]) : "r"(a), "r"(b)), r ^= x[k])
// IMAD.HI.U32
PIPE_KERNEL(k_imad_hi, u32 x[CHAINS], x[k] = a + k,
            asm volatile("mad.hi.u32 %0, %0, %1, %2;" : "+r"(x[k]) : "r"(a), "r"(b)), r ^= x[k])
// IMAD.WIDE.U32, RZ addend: x = lo(x) * hi(x')   (multiplier operands are data dependent so
// ptxas can neither hoist nor share the product)
PIPE_KERNEL(k_imad_wide, u64 x[CHAINS], x[k] = ((u64)a << 32) + b + k,
            asm volatile("{ .reg .u32 lo, hi, l2, h2; mov.b64 {lo,hi}, %0; mov.b64 {l2,h2}, %1; mul.wide.u32 %0, lo, h2; }"
                         : "+l"(x[k]) : "l"(x[(k + 1) % CHAINS])), r ^= (u32)x[k] ^ (u32)(x[k] >> 32))
// IMAD.WIDE.U32 with a 64-bit accumulator: x[k] += lo(x[k+1]) * b
PIPE_KERNEL(k_imad_wide_acc, u64 x[CHAINS], x[k] = ((u64)a << 32) + b + k,
            asm volatile("{ .reg .u32 l2, h2; mov.b64 {l2,h2}, %1; mad.wide.u32 %0, l2, %2, %0; }"
                         : "+l"(x[k]) : "l"(x[(k + 1) % CHAINS]), "r"(b)), r ^= (u32)x[k] ^ (u32)(x[k] >> 32))
// IADD3
PIPE_KERNEL(k_iadd3, u32 x[CHAINS], x[k] = a + k,
            asm volatile("{ .reg .u32 t; add.u32 t, %0, %1; add.u32 %0, t, %2; }" : "+r"(x[k]) : "r"(a), "r"(b)), r ^= x[k])
// LOP3
PIPE_KERNEL(k_lop3, u32 x[CHAINS], x[k] = a + k,
            asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[k]) : "r"(a), "r"(b)), r ^= x[k])
// SHF (funnel shift)
PIPE_KERNEL(k_shf, u32 x[CHAINS], x[k] = a + k,
            asm volatile("shf.l.wrap.b32 %0, %0, %1, 1;" : "+r"(x[k]) : "r"(a)), r ^= x[k])
// mix 1:1  IMAD.WIDE + IADD3 (different chains, so both pipes can run)
PIPE_KERNEL(k_mix_wide_iadd, u64 x[CHAINS]; u32 y[CHAINS], x[k] = ((u64)a << 32) + b + k; y[k] = b + k,
            asm volatile("{ .reg .u32 lo, hi, l2, h2, t; mov.b64 {lo,hi}, %0; mov.b64 {l2,h2}, %2; mul.wide.u32 %0, lo, h2;\n\t"
                         "add.u32 t, %1, %3; add.u32 %1, t, %4; }"
                         : "+l"(x[k]), "+r"(y[k]) : "l"(x[(k + 1) % CHAINS]), "r"(a), "r"(b)), r ^= (u32)x[k] ^ (u32)(x[k] >> 32) ^ y[k])
// mix 1:2  IMAD.WIDE + 2 ALU (IADD3, LOP3)
PIPE_KERNEL(k_mix_wide_2alu, u64 x[CHAINS]; u32 y[CHAINS]; u32 z[CHAINS], x[k] = ((u64)a << 32) + b + k; y[k] = b + k; z[k] = a ^ k,
            asm volatile("{ .reg .u32 lo, hi, l2, h2, t; mov.b64 {lo,hi}, %0; mov.b64 {l2,h2}, %3; mul.wide.u32 %0, lo, h2;\n\t"
                         "add.u32 t, %1, %4; add.u32 %1, t, %5; }\n\tlop3.b32 %2, %2, %4, %5, 0x96;"
                         : "+l"(x[k]), "+r"(y[k]), "+r"(z[k]) : "l"(x[(k + 1) % CHAINS]), "r"(a), "r"(b)), r ^= (u32)x[k] ^ (u32)(x[k] >> 32) ^ y[k] ^ z[k])
// mix 1:1 IMAD.lo + IADD3
PIPE_KERNEL(k_mix_lo_iadd, u32 x[CHAINS]; u32 y[CHAINS], x[k] = a + k; y[k] = b + k,
            asm volatile("mad.lo.u32 %0, %0, %2, %3;\n\t{ .reg .u32 t; add.u32 t, %1, %2; add.u32 %1, t, %3; }"
                         : "+r"(x[k]), "+r"(y[k]) : "r"(a), "r"(b)), r ^= x[k] ^ y[k])
// mix IMAD.WIDE + 32-bit IMAD 1:1 (same pipe?)
PIPE_KERNEL(k_mix_wide_lo, u64 x[CHAINS]; u32 y[CHAINS], x[k] = ((u64)a << 32) + b + k; y[k] = b + k,
            asm volatile("{ .reg .u32 lo, hi, l2, h2; mov.b64 {lo,hi}, %0; mov.b64 {l2,h2}, %2; mul.wide.u32 %0, lo, h2;\n\t"
                         "mad.lo.u32 %1, %1, %3, %4; }"
                         : "+l"(x[k]), "+r"(y[k]) : "l"(x[(k + 1) % CHAINS]), "r"(a), "r"(b)), r ^= (u32)x[k] ^ (u32)(x[k] >> 32) ^ y[k])
// ---------------------------------------------------------------------------
// 63-bit modular multiplication candidates (p = 2^63 - c).  MM/s is what the
// evaluator is made of; see firefly_b200/csrc/ff63.cuh for the shipped version.
// ---------------------------------------------------------------------------
#include "../firefly_b200/csrc/ff63.cuh"

// experimental: "lazy63" variant (inputs < 2^63+2^32, split at bit 63, no final correction)
__device__ __forceinline__ u64 mul_l63_asm(u64 x, u64 y, u32 c) {
  u64 r;
  asm("{\n\t"
      ".reg .u32 x0,x1,y0,y1,w0,w1,w2,w3,a,ml,mh,tl,th,h0,h1,u0,u1,v0,v1,r0,r1,r2,top;\n\t"
      ".reg .u64 t0,m,t3,u,v,ad;\n\t"
      "mov.b64 {x0,x1}, %1;\n\t"
      "mov.b64 {y0,y1}, %2;\n\t"
      "mul.wide.u32 t0, x0, y0;\n\t"
      "mul.wide.u32 m, x0, y1;\n\t"
      "mad.wide.u32 m, x1, y0, m;\n\t"
      "mul.wide.u32 t3, x1, y1;\n\t"
      "mov.b64 {w0,a}, t0;\n\t"
      "mov.b64 {ml,mh}, m;\n\t"
      "mov.b64 {tl,th}, t3;\n\t"
      "add.cc.u32 w1, a, ml;\n\t"
      "addc.cc.u32 w2, tl, mh;\n\t"
      "addc.u32 w3, th, 0;\n\t"
      "shf.l.wrap.b32 h0, w1, w2, 1;\n\t"
      "shf.l.wrap.b32 h1, w2, w3, 1;\n\t"
      "and.b32 w1, w1, 0x7fffffff;\n\t"
      "mul.wide.u32 u, h0, %3;\n\t"
      "mul.wide.u32 v, h1, %3;\n\t"
      "mov.b64 {u0,u1}, u;\n\t"
      "mov.b64 {v0,v1}, v;\n\t"
      "add.cc.u32 r0, w0, u0;\n\t"
      "addc.cc.u32 r1, w1, u1;\n\t"
      "addc.u32 r2, v1, 0;\n\t"
      "add.cc.u32 r1, r1, v0;\n\t"
      "addc.u32 r2, r2, 0;\n\t"
      "shf.l.wrap.b32 top, r1, r2, 1;\n\t"
      "and.b32 r1, r1, 0x7fffffff;\n\t"
      "mov.b64 ad, {r0,r1};\n\t"
      "mad.wide.u32 %0, top, %3, ad;\n\t"
      "}" : "=l"(r) : "l"(x), "l"(y), "r"(c));
  return r;
}

// "b63" family: residues bounded below 2^63 + 2^32 (so both high words are <= 2^31): the two cross
// products are summed inside one IMAD.WIDE (no overflow), fold 1 at 2^64 = d, fold 2 at 2^63 = c with
// no conditional correction.  6 wide + 1 narrow multiplies, 11 adds/logic.  These are the FIRST forms, kept here
// for the comparison: ptxas turns about three of their additions per product into IMAD.X (fma pipe).  What
// ships is ff63::mul_b63 / sqr_b63 (variants 9 and 10 below).
__device__ __forceinline__ u64 mul_b63(u64 x, u64 y, u32 c, u32 d) {
  u64 r;
  asm("{\n\t"
      ".reg .u32 x0,x1,y0,y1,w0,w1,w2,w3,a,ml,mh,tl,th,u0,u1,v0,v1,r0,r1,r2,top,t;\n\t"
      ".reg .u64 t0,m,t3,u,v;\n\t"
      "mov.b64 {x0,x1}, %1;\n\t"
      "mov.b64 {y0,y1}, %2;\n\t"
      "mul.wide.u32 t0, x0, y0;\n\t"
      "mul.wide.u32 m, x0, y1;\n\t"
      "mad.wide.u32 m, x1, y0, m;\n\t"
      "mul.wide.u32 t3, x1, y1;\n\t"
      "mov.b64 {w0,a}, t0;\n\t"
      "mov.b64 {ml,mh}, m;\n\t"
      "mov.b64 {tl,th}, t3;\n\t"
      "add.cc.u32 w1, a, ml;\n\t"
      "addc.cc.u32 w2, tl, mh;\n\t"
      "addc.u32 w3, th, 0;\n\t"
      "mul.wide.u32 u, w2, %4;\n\t"
      "mul.wide.u32 v, w3, %4;\n\t"
      "mov.b64 {u0,u1}, u;\n\t"
      "mov.b64 {v0,v1}, v;\n\t"
      "add.cc.u32 r0, w0, u0;\n\t"
      "addc.cc.u32 r1, w1, u1;\n\t"
      "addc.u32 r2, v1, 0;\n\t"
      "add.cc.u32 r1, r1, v0;\n\t"
      "addc.u32 r2, r2, 0;\n\t"
      "shf.l.wrap.b32 top, r1, r2, 1;\n\t"
      "and.b32 r1, r1, 0x7fffffff;\n\t"
      "mul.lo.u32 t, top, %3;\n\t"
      "add.cc.u32 r0, r0, t;\n\t"
      "addc.u32 r1, r1, 0;\n\t"
      "mov.b64 %0, {r0,r1};\n\t"
      "}" : "=l"(r) : "l"(x), "l"(y), "r"(c), "r"(d));
  return r;
}
// squaring in the same representation: one cross product, doubled by a 64-bit add
__device__ __forceinline__ u64 sqr_b63(u64 x, u32 c, u32 d) {
  u64 r;
  asm("{\n\t"
      ".reg .u32 x0,x1,w0,w1,w2,w3,a,ml,mh,tl,th,u0,u1,v0,v1,r0,r1,r2,top,t;\n\t"
      ".reg .u64 t0,m,t3,u,v;\n\t"
      "mov.b64 {x0,x1}, %1;\n\t"
      "mul.wide.u32 t0, x0, x0;\n\t"
      "mul.wide.u32 m, x0, x1;\n\t"
      "mul.wide.u32 t3, x1, x1;\n\t"
      "mov.b64 {w0,a}, t0;\n\t"
      "mov.b64 {ml,mh}, m;\n\t"
      "mov.b64 {tl,th}, t3;\n\t"
      "add.cc.u32 ml, ml, ml;\n\t"
      "addc.u32 mh, mh, mh;\n\t"
      "add.cc.u32 w1, a, ml;\n\t"
      "addc.cc.u32 w2, tl, mh;\n\t"
      "addc.u32 w3, th, 0;\n\t"
      "mul.wide.u32 u, w2, %3;\n\t"
      "mul.wide.u32 v, w3, %3;\n\t"
      "mov.b64 {u0,u1}, u;\n\t"
      "mov.b64 {v0,v1}, v;\n\t"
      "add.cc.u32 r0, w0, u0;\n\t"
      "addc.cc.u32 r1, w1, u1;\n\t"
      "addc.u32 r2, v1, 0;\n\t"
      "add.cc.u32 r1, r1, v0;\n\t"
      "addc.u32 r2, r2, 0;\n\t"
      "shf.l.wrap.b32 top, r1, r2, 1;\n\t"
      "and.b32 r1, r1, 0x7fffffff;\n\t"
      "mul.lo.u32 t, top, %2;\n\t"
      "add.cc.u32 r0, r0, t;\n\t"
      "addc.u32 r1, r1, 0;\n\t"
      "mov.b64 %0, {r0,r1};\n\t"
      "}" : "=l"(r) : "l"(x), "r"(c), "r"(d));
  return r;
}

template <int K, int VARIANT>
__global__ void __launch_bounds__(256) k_mm(u64* out, const u64* in, u32 c, int trips) {
  const u64 p = 0x8000000000000000ULL - c;
  const u32 d = 2 * c;
  u64 a[K];
  u64 b = in[(threadIdx.x & 31) + 32] % p;
#pragma unroll
  for (int k = 0; k < K; k++) a[k] = (in[threadIdx.x & 31] + 977 * k) % p;
#pragma unroll 1
  for (int t = 0; t < trips; t++) {
#pragma unroll
    for (int i = 0; i < 4; i++) {
#pragma unroll
      for (int k = 0; k < K; k++) {
        if (VARIANT == 0) a[k] = ff63::mul_canon_pm(a[k], b, c, p);
        if (VARIANT == 1) a[k] = ff63::mul_pm(a[k], b, d);
        if (VARIANT == 2) a[k] = ff63::sqr_pm(a[k], d);
        if (VARIANT == 3) a[k] = ff63::mul_pm_dev(a[k], b, d);
        if (VARIANT == 4) a[k] = mul_l63_asm(a[k], b, c);
        if (VARIANT == 5) a[k] = mul_b63(a[k], b, c, d);
        if (VARIANT == 6) a[k] = sqr_b63(a[k], c, d);
        if (VARIANT == 7) a[k] = mul_b63(a[k], a[k], c, d);
        if (VARIANT == 8) a[k] = ff63::sqr_pm_dev(a[k], d);
        if (VARIANT == 9) a[k] = ff63::mul_b63(a[k], b, c, d);        // shipped forms: nothing ptxas can move to the fma pipe
        if (VARIANT == 10) a[k] = ff63::sqr_b63(a[k], c, d);
      }
    }
  }
  u64 r = 0;
#pragma unroll
  for (int k = 0; k < K; k++) r ^= ff63::canon_pm(a[k], c, p);
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

// host replay of thread 0..31 of block 0 for the self-check
static u64 host_mm(const u64* in, int lane, u32 c, int trips, int K, int variant) {
  typedef unsigned __int128 u128;
  const u64 p = 0x8000000000000000ULL - c;
  u64 b = in[lane + 32] % p, r = 0;
  for (int k = 0; k < K; k++) {
    u64 a = (in[lane] + 977ULL * k) % p;
    for (int t = 0; t < trips * 4; t++) a = (variant == 2 || variant == 6 || variant == 7 || variant == 8 || variant == 10) ? (u64)(((u128)a * a) % p) : (u64)(((u128)a * b) % p);
    r ^= a;
  }
  return r;
}

struct Result { std::string name; double ginst_per_s; double per_clk_sm; double ms; };

static double sm_clock_mhz_under_load = 0;

template <typename F>
static double time_ms(F launch, int reps) {
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  launch(); launch();
  CK(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int r = 0; r < reps; r++) {
    CK(cudaEventRecord(e0));
    launch();
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    if (ms < best) best = ms;
  }
  return best;
}

__global__ void k_clock(u64* out, int spin) {
  u64 g0, g1; long long c0, c1;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g0));
  c0 = clock64();
  u32 x = threadIdx.x;
  for (int i = 0; i < spin; i++) asm volatile("mad.lo.u32 %0, %0, %0, %0;" : "+r"(x));
  c1 = clock64();
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g1));
  if (threadIdx.x == 0 && blockIdx.x == 0) { out[0] = g1 - g0; out[1] = (u64)(c1 - c0); out[2] = x; }
}

int main(int argc, char** argv) {
  int dev = 0; CK(cudaSetDevice(dev));
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, dev));
  const int sms = prop.multiProcessorCount;
  const int trips = 2000;
  const int blocks = sms * 8;      // 8 CTAs x 256 thr = 64 warps/SM
  const int threads = 256;
  u32* d_in; u32* d_out; u64* d64_in; u64* d64_out;
  CK(cudaMalloc(&d_in, 64 * 8)); CK(cudaMalloc(&d_out, 4096 * 8));
  CK(cudaMalloc(&d64_out, (size_t)blocks * threads * 8));
  std::vector<u64> h(64);
  for (int i = 0; i < 64; i++) h[i] = 0x9E3779B97F4A7C15ULL * (i + 1) | 1;
  CK(cudaMemcpy(d_in, h.data(), 64 * 8, cudaMemcpyHostToDevice));
  d64_in = (u64*)d_in;

  // effective SM clock under load (globaltimer ns vs clock64)
  u64* d_clk; CK(cudaMalloc(&d_clk, 32));
  k_clock<<<sms * 4, 256>>>(d_clk, 2000000);
  CK(cudaDeviceSynchronize());
  k_clock<<<sms * 4, 256>>>(d_clk, 4000000);
  CK(cudaDeviceSynchronize());
  u64 hc[3]; CK(cudaMemcpy(hc, d_clk, 24, cudaMemcpyDeviceToHost));
  sm_clock_mhz_under_load = (double)hc[1] / (double)hc[0] * 1e3;

  std::vector<Result> res;
  auto run = [&](const char* name, auto kern, double inst_per_thread_trip) {
    double ms = time_ms([&] { kern<<<blocks, threads>>>(d_out, d_in, trips); }, 5);
    CK(cudaGetLastError());
    double total = inst_per_thread_trip * trips * (double)blocks * threads;   // thread-instructions
    double gips = total / (ms * 1e-3) / 1e9;
    res.push_back({name, gips, gips * 1e9 / (sm_clock_mhz_under_load * 1e6) / sms, ms});
  };
  const double per_trip = (double)CHAINS * INNER;
  run("imad_lo", k_imad_lo, per_trip);
  run("imad_hi", k_imad_hi, per_trip);
  run("imad_wide", k_imad_wide, per_trip);
  run("imad_wide_acc", k_imad_wide_acc, per_trip);
  run("iadd3", k_iadd3, per_trip);       // the two PTX adds fuse into one 3-input IADD3
  run("lop3", k_lop3, per_trip);
  run("shf", k_shf, per_trip);
  run("mix_wide_iadd(1w+1a)", k_mix_wide_iadd, per_trip * 2);
  run("mix_wide_2alu(1w+2a)", k_mix_wide_2alu, per_trip * 3);
  run("mix_lo_iadd(1m+1a)", k_mix_lo_iadd, per_trip * 2);
  run("mix_wide_lo(1w+1m)", k_mix_wide_lo, per_trip * 2);

  // modular multiplications / s
  struct MM { std::string name; double gmm; double ms; };
  std::vector<MM> mms;
  const u32 c = 25;
  int n_bad = 0;
  auto run_mm = [&](const char* name, auto kern, int K, int variant) {
    double ms = time_ms([&] { kern<<<blocks, threads>>>(d64_out, d64_in, c, trips); }, 5);
    CK(cudaGetLastError());
    double total = 4.0 * K * trips * (double)blocks * threads;
    mms.push_back({name, total / (ms * 1e-3) / 1e9, ms});
    u64 got[32];
    CK(cudaMemcpy(got, d64_out, sizeof(got), cudaMemcpyDeviceToHost));
    for (int l = 0; l < 32; l += 7) {
      u64 want = host_mm(h.data(), l, c, trips, K, variant);
      if (want != got[l]) { n_bad++; fprintf(stderr, "MISMATCH %s lane %d got %llu want %llu\n", name, l, got[l], want); }
    }
  };
  run_mm("mul_canon_k1", k_mm<1, 0>, 1, 0);
  run_mm("mul_canon_k2", k_mm<2, 0>, 2, 0);
  run_mm("mul_canon_k4", k_mm<4, 0>, 4, 0);
  run_mm("mul_pm_k1", k_mm<1, 1>, 1, 1);
  run_mm("mul_pm_k2", k_mm<2, 1>, 2, 1);
  run_mm("mul_pm_k4", k_mm<4, 1>, 4, 1);
  run_mm("sqr_pm_k1", k_mm<1, 2>, 1, 2);
  run_mm("sqr_pm_k2", k_mm<2, 2>, 2, 2);
  run_mm("sqr_pm_k4", k_mm<4, 2>, 4, 2);
  run_mm("mul_pm_dev_k1", k_mm<1, 3>, 1, 3);
  run_mm("mul_pm_dev_k2", k_mm<2, 3>, 2, 3);
  run_mm("mul_pm_dev_k4", k_mm<4, 3>, 4, 3);
  run_mm("mul_l63_asm_k1", k_mm<1, 4>, 1, 4);
  run_mm("mul_l63_asm_k2", k_mm<2, 4>, 2, 4);
  run_mm("mul_l63_asm_k4", k_mm<4, 4>, 4, 4);
  run_mm("mul_b63_k2", k_mm<2, 5>, 2, 5);
  run_mm("mul_b63_k4", k_mm<4, 5>, 4, 5);
  run_mm("mul_b63_k8", k_mm<8, 5>, 8, 5);
  run_mm("sqr_b63_k4", k_mm<4, 6>, 4, 6);
  run_mm("sqr_b63_k8", k_mm<8, 6>, 8, 6);
  run_mm("sqr_via_mul_b63_k4", k_mm<4, 7>, 4, 7);
  run_mm("sqr_pm_dev_k4", k_mm<4, 8>, 4, 8);
  run_mm("mul_pm_dev_k8", k_mm<8, 3>, 8, 3);
  run_mm("mul_b63_v2_k4", k_mm<4, 9>, 4, 9);
  run_mm("mul_b63_v2_k8", k_mm<8, 9>, 8, 9);
  run_mm("sqr_b63_v2_k4", k_mm<4, 10>, 4, 10);
  run_mm("sqr_b63_v2_k8", k_mm<8, 10>, 8, 10);

  // occupancy sweep of the shipped multiplication: resident warps per SM limited through dynamic shared memory
  struct Occ { int k, warps; double gmm; };
  std::vector<Occ> occ;
  {
    auto sweep = [&](auto kern, int K) {
      CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
      for (int bps : {1, 2, 3, 4, 6, 8}) {           // blocks of 256 threads per SM
        size_t smem = (size_t)(227 * 1024 / bps) - 2048;
        if (smem > 200 * 1024) smem = 200 * 1024;
        const int nb = sms * bps * 4;
        double ms = time_ms([&] { kern<<<nb, threads, smem>>>(d64_out, d64_in, c, trips / 4); }, 3);
        CK(cudaGetLastError());
        double total = 4.0 * K * (trips / 4) * (double)nb * threads;
        occ.push_back({K, bps * 8, total / (ms * 1e-3) / 1e9});
      }
    };
    u64* big; CK(cudaMalloc(&big, (size_t)sms * 8 * 4 * threads * 8));
    u64* keep = d64_out; d64_out = big;
    sweep(k_mm<4, 5>, 4);
    sweep(k_mm<8, 5>, 8);
    d64_out = keep;
  }

  printf("{\"gpu\": \"%s\", \"sms\": %d, \"sm_clock_mhz_under_load\": %.1f, \"blocks\": %d, \"threads\": %d,\n \"pipes\": {",
         prop.name, sms, sm_clock_mhz_under_load, blocks, threads);
  for (size_t i = 0; i < res.size(); i++)
    printf("%s\n  \"%s\": {\"gthread_inst_per_s\": %.1f, \"thread_inst_per_clk_per_sm\": %.2f, \"ms\": %.3f}",
           i ? "," : "", res[i].name.c_str(), res[i].ginst_per_s, res[i].per_clk_sm, res[i].ms);
  printf("},\n \"modmul\": {");
  for (size_t i = 0; i < mms.size(); i++)
    printf("%s\n  \"%s\": {\"gmm_per_s\": %.2f, \"ms\": %.3f}", i ? "," : "", mms[i].name.c_str(), mms[i].gmm, mms[i].ms);
  printf("},\n \"mul_b63_by_occupancy\": [");
  for (size_t i = 0; i < occ.size(); i++)
    printf("%s{\"chains_per_thread\": %d, \"warps_per_sm\": %d, \"gmm_per_s\": %.1f}", i ? ", " : "", occ[i].k, occ[i].warps, occ[i].gmm);
  printf("],\n \"selfcheck_mismatches\": %d}\n", n_bad);
  return n_bad != 0;
}
