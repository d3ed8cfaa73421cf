// kernels.cu -- the sm_100a tape interpreter: batched black-box probe evaluation mod p.
//
// Replaces the reference's hot loop, ShuntingYardParser::evaluate_pre<FFInt / FFIntVec<N>>
// (reference: source/include/firefly/ShuntingYardParser.hpp:365-479) together with the FFInt
// operators it calls (source/FFInt.cpp:130-192) for ALL probe points of a batch at once.
//
// Mapping to the machine
//   * one probe point per lane, K points per thread (independent multiply chains for ILP); at K = 8 the
//     full build takes 80 registers (three resident 256-thread blocks per SM), the LEAN build -- tapes whose
//     instructions all have a dense dispatch code, i.e. every benchmark function -- 64 (four blocks);
//   * the accumulator of the tape machine (tape.hpp) is K residues of two 32-bit registers each;
//     intermediate values are BOUNDED lazy residues (below 2^63 + 2^32, ff63.cuh) and are made
//     canonical only where the reference's semantics need it: zero tests of divisors, lane
//     exponents, and stores;
//   * the tape is warp-uniform: every branch of the interpreter is uniform, instructions are
//     fetched with one 128-bit load one instruction ahead of execution; the frequent instructions
//     are cases of one switch over a dense code, the rare ones go through generic_op on copies;
//   * spill and temp slots live in shared memory, laid out [slot][k][thread] so that a warp's
//     access is one conflict-free 256-byte row;
//   * probe coordinates are read as 16-byte vectors, [var][point] layout, fully coalesced
//     (DRAM traffic equals the algorithmic 8 * (n_vars + n_out) bytes per probe, profiles/);
//     results are written [function][point], 16-byte vectors, fully coalesced;
//   * divisions: ONE Fermat inversion per thread block.  Every lane's divisor enters a
//     three-level product tree (K values per thread -> 32 lanes by shuffle scans -> the
//     block's warps through shared memory); lane w of warp 0 inverts the running product of
//     warps 0..w, so the lanes of the inverting warp all do useful, different work.
//     Zero divisors are replaced by 1 in the tree and produce 0 (reference: a/0 = 0).
#include "kernels.cuh"

#include <algorithm>

#include "ff63.cuh"

namespace ffgpu {

using ff63::u32;
using ff63::u64;

namespace {

// Two arithmetics, chosen per launch (EvalParams::generic):
//   Field    p = 2^63 - c, c < 2^15 -- every prime of FireFly's table (reference: source/ReconstHelper.cpp:24-125).
//            Bounded lazy residues, folds by c (ff63.cuh, "b63" family).  The tuned path.
//   FieldB   any odd prime below 2^63 (FFInt::set_new_prime takes any, reference: source/FFInt.cpp:344-353; its
//            mod_mul is generic, :356-369).  Canonical residues, Barrett reduction with a normalised divisor
//            (ff63::mul_barrett): about three times the multiply instructions of the tuned path.
struct Field {
  u64 p;
  u32 c, d;
};
struct FieldB {
  u64 p, pn, pinv;
  u32 k;
};
template <class FT>
__device__ __forceinline__ FT make_field(const EvalParams& prm);
template <>
__device__ __forceinline__ Field make_field<Field>(const EvalParams& prm) { return Field{prm.p, prm.c, prm.d}; }
template <>
__device__ __forceinline__ FieldB make_field<FieldB>(const EvalParams& prm) { return FieldB{prm.p, prm.p << prm.pshift, prm.pinv, prm.pshift}; }

// One residue = two 32-bit words in two independent registers (no aligned pair: see ff63.cuh, word-level
// forms).  Values are BOUNDED lazy residues (ff63.cuh, "b63" family): below 2^63 + 2^32; canonical inputs
// qualify, every result qualifies again.
struct R {
  u32 lo, hi;
};
__device__ __forceinline__ R mkR(u64 v) { return R{(u32)v, (u32)(v >> 32)}; }
__device__ __forceinline__ u64 toU(R a) { return ((u64)a.hi << 32) | a.lo; }

__device__ __forceinline__ R fmul(R a, R b, const Field& F) {
  R r;
  ff63::mul_b63_w(r.lo, r.hi, a.lo, a.hi, b.lo, b.hi, F.c, F.d);
  return r;
}
__device__ __forceinline__ R fsqr(R a, const Field& F) {
  R r;
  ff63::sqr_b63_w(r.lo, r.hi, a.lo, a.hi, F.c, F.d);
  return r;
}
__device__ __forceinline__ R fadd(R a, R b, const Field& F) {
  R r;
  ff63::add_b63_w(r.lo, r.hi, a.lo, a.hi, b.lo, b.hi, F.c);
  return r;
}
__device__ __forceinline__ R fsub(R a, R b, const Field& F) {
  R r;
  ff63::sub_b63_w(r.lo, r.hi, a.lo, a.hi, b.lo, b.hi, F.c, F.d);
  return r;
}
__device__ __forceinline__ R fneg(R a, const Field& F) {
  R r;
  ff63::neg_b63_w(r.lo, r.hi, a.lo, a.hi, F.c, F.d);
  return r;
}
__device__ __forceinline__ u64 fcanon(R a, const Field& F) { return ff63::canon_pm(toU(a), F.c, F.p); }
// what a caller hands in (documented: canonical residues): the tuned path computes with anything below 2^63 + 2^32
__device__ __forceinline__ R fin(R a, const Field&) { return a; }
// ---- generic modulus: canonical residues throughout
__device__ __forceinline__ ff63::Barrett barrett_of(const FieldB& F) { return ff63::Barrett{F.p, F.pn, F.pinv, F.k}; }
__device__ __forceinline__ R fmul(R a, R b, const FieldB& F) { return mkR(ff63::mul_barrett(toU(a), toU(b), barrett_of(F))); }
__device__ __forceinline__ R fsqr(R a, const FieldB& F) { return mkR(ff63::mul_barrett(toU(a), toU(a), barrett_of(F))); }
__device__ __forceinline__ R fadd(R a, R b, const FieldB& F) { return mkR(ff63::add(toU(a), toU(b), F.p)); }
__device__ __forceinline__ R fsub(R a, R b, const FieldB& F) { return mkR(ff63::sub(toU(a), toU(b), F.p)); }
__device__ __forceinline__ R fneg(R a, const FieldB& F) { return mkR(ff63::neg(toU(a), F.p)); }
__device__ __forceinline__ u64 fcanon(R a, const FieldB&) { return toU(a); }
__device__ __forceinline__ R fin(R a, const FieldB& F) {
  u64 v = toU(a);
  if (v >= F.p) v %= F.p;        // a caller value that is not a residue: reduce it, never compute on with it
  return mkR(v);
}
// a constant the compiler may not hoist or speculate (plain broadcasts of an immediate into all accumulators
// were moved above the interpreter's dispatch, 16 register moves per tape instruction)
__device__ __forceinline__ R opaque(u64 v) {
  R r;
  asm volatile("mov.b64 {%0,%1}, %2;" : "=r"(r.lo), "=r"(r.hi) : "l"(v));
  return r;
}
__device__ __forceinline__ R shfl_up(R a, unsigned off) {
  return R{__shfl_up_sync(0xffffffffu, a.lo, off), __shfl_up_sync(0xffffffffu, a.hi, off)};
}
__device__ __forceinline__ R shfl_down(R a, unsigned off) {
  return R{__shfl_down_sync(0xffffffffu, a.lo, off), __shfl_down_sync(0xffffffffu, a.hi, off)};
}
__device__ __forceinline__ R shfl_idx(R a, int src) {
  return R{__shfl_sync(0xffffffffu, a.lo, src), __shfl_sync(0xffffffffu, a.hi, src)};
}

// x^e, e uniform across the warp, left-to-right binary; x^0 = 1 (reference mod_pow, FFInt.cpp:371-383)
template <int K, class FT>
__device__ __forceinline__ void pow_uniform(R (&a)[K], u64 e, const FT& F) {
  if (e <= 1) {
    if (e == 0) {
#pragma unroll
      for (int k = 0; k < K; k++) a[k] = opaque(1);
    }
    return;
  }
  // The running power stays in the accumulator's own registers and the base is an explicit copy that the
  // compiler may not move: with the roles the other way round (or a plain C++ copy) ptxas double-buffered
  // ALL accumulators at EVERY tape instruction -- 16 IMAD.MOV per dispatch on the multiply pipe.
  R base[K];
#pragma unroll
  for (int k = 0; k < K; k++)
    asm volatile("mov.b32 %0, %2;\n\tmov.b32 %1, %3;" : "=r"(base[k].lo), "=r"(base[k].hi) : "r"(a[k].lo), "r"(a[k].hi));
  int bit = 62 - __clzll((long long)e);          // index below the leading one
#pragma unroll 1
  for (; bit >= 0; --bit) {
#pragma unroll
    for (int k = 0; k < K; k++) a[k] = fsqr(a[k], F);
    if ((e >> bit) & 1) {
#pragma unroll
      for (int k = 0; k < K; k++) a[k] = fmul(a[k], base[k], F);
    }
  }
}

// r = base^e for e >= 1, out of place: the base keeps its registers (it was just loaded), the running power is
// born in r's -- no copies at all
template <int K, class FT>
__device__ __forceinline__ void pow_from(R (&r)[K], const R (&base)[K], u32 e, const FT& F) {
  if (e == 1) {
#pragma unroll
    for (int k = 0; k < K; k++)
      asm volatile("mov.b32 %0, %2;\n\tmov.b32 %1, %3;" : "=r"(r[k].lo), "=r"(r[k].hi) : "r"(base[k].lo), "r"(base[k].hi));
    return;
  }
  int bit = 30 - __clz((int)e);                  // index below the leading one, >= 0 here
#pragma unroll
  for (int k = 0; k < K; k++) r[k] = fsqr(base[k], F);
  if ((e >> bit) & 1) {
#pragma unroll
    for (int k = 0; k < K; k++) r[k] = fmul(r[k], base[k], F);
  }
#pragma unroll 1
  for (--bit; bit >= 0; --bit) {
#pragma unroll
    for (int k = 0; k < K; k++) r[k] = fsqr(r[k], F);
    if ((e >> bit) & 1) {
#pragma unroll
      for (int k = 0; k < K; k++) r[k] = fmul(r[k], base[k], F);
    }
  }
}

template <class FT>
__device__ __forceinline__ R sqr_n(R x, int n, const FT& F) {
#pragma unroll 1
  for (int i = 0; i < n; i++) x = fsqr(x, F);
  return x;
}

// x^(p-2) for p = 2^63 - c:  p - 2 = (2^47 - 1) * 2^16 + (2^16 - c - 2).
// 2^47 - 1 by the addition chain 1,2,3,5,10,20,23,46,47 on the run length; then 16 squarings
// interleaved with the bits of the 16-bit tail.  62 squarings + ~16 multiplications.  0 -> 0.
__device__ __noinline__ R inv_fermat(R x, Field F) {
  const R x1 = x;
  const R x2 = fmul(fsqr(x1, F), x1, F);
  const R x3 = fmul(fsqr(x2, F), x1, F);
  const R x5 = fmul(sqr_n(x3, 2, F), x2, F);
  const R x10 = fmul(sqr_n(x5, 5, F), x5, F);
  const R x20 = fmul(sqr_n(x10, 10, F), x10, F);
  const R x23 = fmul(sqr_n(x20, 3, F), x3, F);
  const R x46 = fmul(sqr_n(x23, 23, F), x23, F);
  R r = fmul(fsqr(x46, F), x1, F);               // x^(2^47 - 1)
  const u32 tail = 65536u - F.c - 2u;            // c < 2^15: 16 bits
#pragma unroll 1
  for (int bit = 15; bit >= 0; --bit) {
    r = fsqr(r, F);
    if ((tail >> bit) & 1) r = fmul(r, x1, F);
  }
  return r;
}

// x^(p-2) for any odd prime p: left-to-right binary ladder over the bits of p - 2.  0 -> 0.
__device__ __noinline__ R inv_fermat(R x, FieldB F) {
  const u64 e = F.p - 2;
  if (e == 1) return x;                           // p = 3
  R r = x;
#pragma unroll 1
  for (int bit = 62 - __clzll((long long)e); bit >= 0; --bit) {
    r = fsqr(r, F);
    if ((e >> bit) & 1) r = fmul(r, x, F);
  }
  return r;
}

// per-lane power with a lane-dependent exponent (divergent; only for `x^y` style input)
template <class FT>
__device__ __noinline__ R pow_lane(R b, u64 e, FT F) {
  R r = mkR(1);
#pragma unroll 1
  while (e) {
    if (e & 1) r = fmul(r, b, F);
    b = fsqr(b, F);
    e >>= 1;
  }
  return r;
}
// FFInt::pow(const FFInt&) per lane (FFInt.cpp:155-179)
template <class FT>
__device__ __forceinline__ R pow_ff_lane(R b, R e_lazy, const FT& F) {
  const u64 e = fcanon(e_lazy, F);
  if (e == 2) return fsqr(b, F);
  if (e < (F.p >> 1)) return pow_lane(b, e, F);
  return pow_lane(inv_fermat(b, F), F.p - e - 1, F);
}
// FFInt::pow<int>(to_neg_int()) per lane (FFInt.hpp:149-174, FFInt.cpp:226-228)
template <class FT>
__device__ __forceinline__ R pow_negint_lane(R b, R e_lazy, const FT& F) {
  const int e = (int)(0u - (u32)fcanon(e_lazy, F));
  if (e == 2) return fsqr(b, F);
  if (e >= 0) return pow_lane(b, (u64)e, F);
  return pow_lane(inv_fermat(b, F), (u64)(-(long long)e), F);
}

// ---- block-batched inversion -----------------------------------------------------------------
// a[k] <- 1/a[k] (0 -> 0) for the K values of every thread of the block.  scratch: 64 u64 of
// shared memory.  Must be reached by all threads of the block (the tape is block uniform).
template <int K, int THREADS, class FT>
__device__ __forceinline__ void block_inverse(R (&a)[K], const FT& F, u64* scratch) {
  constexpr int W = THREADS / 32;
  const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  bool z[K];
  R e[K], pre[K];
#pragma unroll
  for (int k = 0; k < K; k++) {
    // the canonical value goes into the tree: a lane holding something out of bounds (a raw caller value)
    // can then only spoil its own result, never the inverses of the other probes of the block
    const u64 cv = fcanon(a[k], F);
    z[k] = cv == 0;
    e[k] = z[k] ? mkR(1) : mkR(cv);
  }
  pre[0] = e[0];
#pragma unroll
  for (int k = 1; k < K; k++) pre[k] = fmul(pre[k - 1], e[k], F);
  const R t = pre[K - 1];
  // inclusive prefix and inclusive suffix products over the lanes of the warp
  R s = t, u = t;
#pragma unroll
  for (int off = 1; off < 32; off <<= 1) {
    const R sm = fmul(s, shfl_up(s, off), F);
    const R um = fmul(u, shfl_down(u, off), F);
    if (lane >= (unsigned)off) s = sm;
    if (lane + off < 32) u = um;
  }
  R sx = shfl_up(s, 1);        // product of the lanes before this one
  R ux = shfl_down(u, 1);      // product of the lanes after this one
  if (lane == 0) sx = mkR(1);
  if (lane == 31) ux = mkR(1);
  R inv_upto, before;          // 1 / (warps 0..warp), product of warps 0..warp-1
  if (W > 1) {
    if (lane == 31) scratch[warp] = toU(s);
    __syncthreads();
    if (warp == 0) {
      R S = lane < W ? mkR(scratch[lane]) : mkR(1);
#pragma unroll
      for (int off = 1; off < W; off <<= 1) {
        const R sm = fmul(S, shfl_up(S, off), F);
        if (lane >= (unsigned)off) S = sm;
      }
      const R IS = inv_fermat(S, F);
      R Sx = shfl_up(S, 1);
      if (lane == 0) Sx = mkR(1);
      scratch[lane] = toU(Sx);
      scratch[32 + lane] = toU(IS);
    }
    __syncthreads();
    before = mkR(scratch[warp]);
    inv_upto = mkR(scratch[32 + warp]);
  } else {
    inv_upto = inv_fermat(shfl_idx(s, 31), F);
    before = mkR(1);
  }
  // 1/t = 1/(everything up to the end of this warp) * (warps before) * (lanes before) * (lanes after)
  R J = fmul(fmul(inv_upto, ux, F), fmul(before, sx, F), F);
#pragma unroll
  for (int k = K - 1; k >= 1; --k) {
    const R q = fmul(J, pre[k - 1], F);
    a[k] = z[k] ? mkR(0) : q;
    J = fmul(J, e[k], F);
  }
  a[0] = z[0] ? mkR(0) : J;
}

// ---- the interpreter -------------------------------------------------------------------------
union InstrWord {
  uint4 v;
  Instr i;
};

// Per-thread view of the batch: where this thread's probes are, its column of the shared-memory slots.
template <int K, int THREADS, bool VEC, class FT>
struct Lane {
  static constexpr u32 STEP = VEC ? 2 * THREADS : THREADS;   // distance between this thread's accesses
  static constexpr int NACC = VEC ? K / 2 : K;               // accesses per variable
  const u64* vbase;      // values + first point of the tile
  u64* obase;            // out + first point of the tile
  u64 in_stride, out_stride;
  uint2* stk;            // [slot][k][thread], this thread's column
  u64* scratch;          // 64 words for block_inverse
  u32 off[NACC];         // tile-relative offsets of this thread's accesses, clamped into the batch
  u32 own, left, tmp_top;
  const u64* tbase;      // shared rows (variables n_base, n_base + 1, ...) + first point of the tile
  u64 t_stride;
  u32 n_base;
  FT F;

  // Points owned by this thread.  VEC: pairs (2t, 2t+1) inside rows of 2*THREADS points, so that every
  // load/store is one 16-byte access per lane and a warp covers 512 contiguous bytes.  Lanes beyond the end
  // of the batch recompute the last valid point and never store: the hot loop has no bounds checks.
  __device__ __forceinline__ void init(const EvalParams& prm, u64* smem) {
    const u64 tile = (u64)blockIdx.x * (u64)(THREADS * K);
    const u64 left64 = prm.n_points - tile;               // > 0: the grid covers exactly the batch
    left = left64 < (u64)(THREADS * K) ? (u32)left64 : (u32)(THREADS * K);
    own = VEC ? 2 * threadIdx.x : threadIdx.x;
#pragma unroll
    for (int j = 0; j < NACC; j++) {
      const u32 o = own + j * STEP, last = left - (VEC ? 2 : 1);
      off[j] = o < last ? o : last;
    }
    vbase = reinterpret_cast<const u64*>(prm.values) + tile;
    obase = reinterpret_cast<u64*>(prm.out) + tile;
    in_stride = prm.in_stride;
    out_stride = prm.out_stride;
    tbase = reinterpret_cast<const u64*>(prm.table) + tile;
    t_stride = prm.table_stride;
    n_base = prm.n_base;
    scratch = smem;
    stk = reinterpret_cast<uint2*>(smem + 64) + threadIdx.x;
    tmp_top = prm.n_slots - 1;
    F = make_field<FT>(prm);
  }
  __device__ __forceinline__ void load_var(u32 var, R (&b)[K]) const {
    const u64* row = var < n_base ? vbase + (u64)var * in_stride : tbase + (u64)(var - n_base) * t_stride;
    if (VEC) {
#pragma unroll
      for (int j = 0; j < NACC; j++) {
        // n_points is even and rows are 16-byte aligned on this path
        const uint4 v = __ldg(reinterpret_cast<const uint4*>(row + off[j]));
        b[2 * j] = fin(R{v.x, v.y}, F);
        b[2 * j + 1] = fin(R{v.z, v.w}, F);
      }
    } else {
#pragma unroll
      for (int j = 0; j < NACC; j++) {
        const uint2 v = __ldg(reinterpret_cast<const uint2*>(row + off[j]));
        b[j] = fin(R{v.x, v.y}, F);
      }
    }
  }
  __device__ __forceinline__ bool in_range(int j) const { return own + j * STEP < left; }
  __device__ __forceinline__ void slot_put(const R (&a)[K], u32 slot) const {
    uint2* t = stk + (size_t)slot * (K * THREADS);
#pragma unroll
    for (int k = 0; k < K; k++) t[k * THREADS] = make_uint2(a[k].lo, a[k].hi);
  }
  __device__ __forceinline__ void slot_get(R (&b)[K], u32 slot) const {
    const uint2* t = stk + (size_t)slot * (K * THREADS);
#pragma unroll
    for (int k = 0; k < K; k++) {
      const uint2 v = t[k * THREADS];
      b[k] = R{v.x, v.y};
    }
  }
  // spill slots count up from 0, temp slots (shared sub-expressions, co-evaluated sums) down from the last slot
  __device__ __forceinline__ void tmp_put(const R (&a)[K], u32 slot) const { slot_put(a, tmp_top - slot); }
  __device__ __forceinline__ void tmp_get(R (&b)[K], u32 slot) const { slot_get(b, tmp_top - slot); }
  __device__ __forceinline__ void store_row(u32 out_slot, const R (&a)[K]) const {
    u64* row = obase + (u64)out_slot * out_stride;
    if (VEC) {
#pragma unroll
      for (int j = 0; j < NACC; j++)
        if (in_range(j))
          *reinterpret_cast<ulonglong2*>(row + off[j]) = make_ulonglong2(fcanon(a[2 * j], F), fcanon(a[2 * j + 1], F));
    } else {
#pragma unroll
      for (int j = 0; j < NACC; j++)
        if (in_range(j)) row[off[j]] = fcanon(a[j], F);
    }
  }
};

#ifndef FFGPU_GENERIC_INLINE
#define FFGPU_GENERIC_INLINE __forceinline__
#endif
// The uncommon instructions (lane exponents, negated variables, negative literal powers, ...), out of line:
// the accumulators travel through local memory for these, which keeps them out of the register allocation
// of the hot loop (inlined, they made ptxas copy all accumulators at every tape instruction).
template <int K, int THREADS, bool VEC, class FT>
__device__ FFGPU_GENERIC_INLINE u32 generic_op(const Lane<K, THREADS, VEC, FT>& L, u32 opword, u32 arg, u64 imm, R (&acc)[K],
                                       u32 sp) {
  struct {
    u32 op, src, flags, arg;
    u64 imm;
  } ins = {opword & 0xffu, (opword >> 8) & 0xffu, (opword >> 16) & 0xffu, arg, imm};
  const FT F = L.F;
  R b[K];
  switch (ins.src) {
    case SRC_POP: L.slot_get(b, --sp); break;
    case SRC_VAR: L.load_var(ins.arg, b); break;
    case SRC_NEGVAR:
      L.load_var(ins.arg, b);
#pragma unroll
      for (int k = 0; k < K; k++) b[k] = fneg(b[k], F);
      break;
    case SRC_CONST:
#pragma unroll
      for (int k = 0; k < K; k++) b[k] = mkR(ins.imm);
      break;
    case SRC_TMP: L.tmp_get(b, ins.arg); break;
    default:
#pragma unroll
      for (int k = 0; k < K; k++) b[k] = mkR(0);
      break;
  }
  switch (ins.op) {
    case OP_LOAD:
      if (ins.flags & F_PUSH) L.slot_put(acc, sp++);
#pragma unroll
      for (int k = 0; k < K; k++) acc[k] = b[k];
      break;
    case OP_ADD:
#pragma unroll
      for (int k = 0; k < K; k++) acc[k] = fadd(acc[k], b[k], F);
      break;
    case OP_SUB:
#pragma unroll
      for (int k = 0; k < K; k++) acc[k] = (ins.flags & F_SWAP) ? fsub(b[k], acc[k], F) : fsub(acc[k], b[k], F);
      break;
    case OP_MUL:
#pragma unroll
      for (int k = 0; k < K; k++) acc[k] = fmul(acc[k], b[k], F);
      break;
    case OP_SQR:
#pragma unroll
      for (int k = 0; k < K; k++) acc[k] = fsqr(acc[k], F);
      break;
    case OP_SAVE:
      L.tmp_put(acc, ins.arg);
      break;
    case OP_ACCUM:
      if (ins.flags & F_INIT) {
#pragma unroll
        for (int k = 0; k < K; k++) b[k] = (ins.flags & F_SWAP) ? fneg(acc[k], F) : acc[k];
      } else {
        L.tmp_get(b, ins.arg);
#pragma unroll
        for (int k = 0; k < K; k++) b[k] = (ins.flags & F_SWAP) ? fsub(b[k], acc[k], F) : fadd(b[k], acc[k], F);
      }
      L.tmp_put(b, ins.arg);
      break;
    case OP_NEG:
#pragma unroll
      for (int k = 0; k < K; k++) acc[k] = fneg(acc[k], F);
      break;
    case OP_INV:
      block_inverse<K, THREADS>(acc, F, L.scratch);
      break;
    case OP_POWU: {
      // FFInt::pow(const FFInt&) with a lane-uniform exponent residue (FFInt.cpp:155-179)
      const u64 e = ins.imm;
      if (e == 2) {
#pragma unroll
        for (int k = 0; k < K; k++) acc[k] = fsqr(acc[k], F);
      } else if (e < (F.p >> 1)) {
        pow_uniform<K>(acc, e, F);
      } else {
        block_inverse<K, THREADS>(acc, F, L.scratch);
        pow_uniform<K>(acc, F.p - e - 1, F);
      }
      break;
    }
    case OP_NPOWU: {
      // FFInt::pow<int>(a.to_neg_int()) with a lane-uniform exponent (FFInt.hpp:149-174)
      const int e = (int)(0u - (u32)ins.imm);
      if (e == 2) {
#pragma unroll
        for (int k = 0; k < K; k++) acc[k] = fsqr(acc[k], F);
      } else if (e >= 0) {
        pow_uniform<K>(acc, (u64)e, F);
      } else {
        block_inverse<K, THREADS>(acc, F, L.scratch);
        pow_uniform<K>(acc, (u64)(-(long long)e), F);
      }
      break;
    }
    case OP_POWL:
#pragma unroll
      for (int k = 0; k < K; k++)
        acc[k] = (ins.flags & F_SWAP) ? pow_ff_lane(b[k], acc[k], F) : pow_ff_lane(acc[k], b[k], F);
      break;
    case OP_NPOWL:
#pragma unroll
      for (int k = 0; k < K; k++)
        acc[k] = (ins.flags & F_SWAP) ? pow_negint_lane(b[k], acc[k], F) : pow_negint_lane(acc[k], b[k], F);
      break;
    case OP_STORE:
      L.store_row(ins.arg, acc);
      break;
    default:
      asm volatile("trap;");     // an instruction without code: never compute on with garbage
      break;
  }
  if (ins.flags & F_NEGR) {
#pragma unroll
    for (int k = 0; k < K; k++) acc[k] = fneg(acc[k], F);
  }
  return sp;
}

// Resident blocks per SM the kernel is compiled for (= its register budget).  At 8 probes per thread the
// interpreter wants ~100 registers: below that ptxas spills the instruction words or double-buffers the
// accumulators at every dispatch, which costs more than the lost occupancy.
// LEAN kernels leave out the generic instructions (tapes that need none are the rule: all benchmark functions)
// and are compiled for 64 registers at 8 probes per thread -- four resident 256-thread blocks instead of three;
// measured +2 % (f1) to +8 % (f3, f4) when shared memory allows the fourth block, nothing to gain otherwise.
template <int K, int THREADS, bool LEAN>
constexpr int min_blocks() {
  return (K == 8 && !LEAN) ? 768 / THREADS : 1024 / THREADS;
}

template <int K, int THREADS, bool VEC, bool LEAN, class FT>
__global__ void __launch_bounds__(THREADS, min_blocks<K, THREADS, LEAN>()) eval_tape_kernel(const EvalParams prm) {
  extern __shared__ __align__(16) u64 smem[];
  Lane<K, THREADS, VEC, FT> L;
  L.init(prm, smem);
  const FT F = L.F;

  // Ask L2 for this tile's variable rows before the tape starts: the interpreter loads a row where an instruction
  // uses it, so the first touch of every variable would otherwise pay a full DRAM round trip one after the other
  // (+2 % on f3 / f4 and the 2000-function collection; with twenty rows as in f1 the requests cost more than they
  // save: profiles/r02_kbench_prefetch_l2.log against r02_kbench_block_sizes.log, same box).
  if (K == 8 && prm.n_vars <= 8u) {
#pragma unroll 1
    for (u32 v = 0; v < prm.n_vars; v++) {
      const u64* row = L.vbase + (u64)v * L.in_stride;
#pragma unroll
      for (int j = 0; j < Lane<K, THREADS, VEC, FT>::NACC; j++) asm volatile("prefetch.global.L2 [%0];" ::"l"(row + L.off[j]));
    }
  }
  R acc[K];
#pragma unroll
  for (int k = 0; k < K; k++) acc[k] = opaque(0);
  u32 sp = 0;                                // spill-stack pointer (uniform)

  // blockIdx.y selects a chunk of the tape (a run of whole functions): many-function collections
  // evaluated at few points still fill the chip
  const uint4* tape = reinterpret_cast<const uint4*>(prm.tape);
  u32 pc = 0, pc_end = prm.n_instr;
  if (prm.chunk_bounds) {
    pc = __ldg(prm.chunk_bounds + blockIdx.y);
    pc_end = __ldg(prm.chunk_bounds + blockIdx.y + 1);
  }
  InstrWord cur, nxt;
  cur.v = __ldg(tape + pc);
  for (; pc < pc_end;) {
    nxt.v = __ldg(tape + pc + 1);            // the tape ends in two HALTs: always in bounds
    const Instr ins = cur.i;
    R b[K];
    switch (ins.fast) {
      // ---- the common (op, operand) combinations, each working on ACC in place -----------------
      case FX_PUSH_LOAD_VAR:
        L.slot_put(acc, sp++);
      case FX_LOAD_VAR:
        L.load_var(ins.arg, acc);
        break;
      case FX_PUSH_LOAD_CONST:
        L.slot_put(acc, sp++);
      case FX_LOAD_CONST:
#pragma unroll
        for (int k = 0; k < K; k++) acc[k] = opaque(ins.imm);
        break;
      case FX_ADD_VAR:
        L.load_var(ins.arg, b);
#pragma unroll
        for (int k = 0; k < K; k++) acc[k] = fadd(acc[k], b[k], F);
        break;
      case FX_ADD_CONST: {
        const R c = mkR(ins.imm);
#pragma unroll
        for (int k = 0; k < K; k++) acc[k] = fadd(acc[k], c, F);
        break;
      }
      case FX_ADD_POP:
        L.slot_get(b, --sp);
#pragma unroll
        for (int k = 0; k < K; k++) acc[k] = fadd(acc[k], b[k], F);
        break;
      case FX_SUB_VAR:
        L.load_var(ins.arg, b);
#pragma unroll
        for (int k = 0; k < K; k++) acc[k] = fsub(acc[k], b[k], F);
        break;
      case FX_RSUB_VAR:
        L.load_var(ins.arg, b);
#pragma unroll
        for (int k = 0; k < K; k++) acc[k] = fsub(b[k], acc[k], F);
        break;
      case FX_RSUB_CONST: {
        const R c = mkR(ins.imm);
#pragma unroll
        for (int k = 0; k < K; k++) acc[k] = fsub(c, acc[k], F);
        break;
      }
      case FX_SUB_POP:
        L.slot_get(b, --sp);
#pragma unroll
        for (int k = 0; k < K; k++) acc[k] = fsub(acc[k], b[k], F);
        break;
      case FX_RSUB_POP:
        L.slot_get(b, --sp);
#pragma unroll
        for (int k = 0; k < K; k++) acc[k] = fsub(b[k], acc[k], F);
        break;
      case FX_MUL_VAR:
        L.load_var(ins.arg, b);
#pragma unroll
        for (int k = 0; k < K; k++) acc[k] = fmul(acc[k], b[k], F);
        break;
      case FX_MUL_CONST: {
        const R c = mkR(ins.imm);
#pragma unroll
        for (int k = 0; k < K; k++) acc[k] = fmul(acc[k], c, F);
        break;
      }
      case FX_MUL_POP:
        L.slot_get(b, --sp);
#pragma unroll
        for (int k = 0; k < K; k++) acc[k] = fmul(acc[k], b[k], F);
        break;
      case FX_SQR:
#pragma unroll
        for (int k = 0; k < K; k++) acc[k] = fsqr(acc[k], F);
        break;
      case FX_POWU:
        // a literal power below p/2 (the compiler gives everything else the generic code)
        pow_uniform<K>(acc, ins.imm, F);
        break;
      case FX_INV:
        block_inverse<K, THREADS>(acc, F, L.scratch);
        break;
      case FX_SAVE:
      case FX_ACCUM_INIT:
        L.tmp_put(acc, ins.arg);
        break;
      case FX_ACCUM_ADD:
        L.tmp_get(b, ins.arg);
#pragma unroll
        for (int k = 0; k < K; k++) b[k] = fadd(b[k], acc[k], F);
        L.tmp_put(b, ins.arg);
        break;
      case FX_ACCUM_SUB:
        L.tmp_get(b, ins.arg);
#pragma unroll
        for (int k = 0; k < K; k++) b[k] = fsub(b[k], acc[k], F);
        L.tmp_put(b, ins.arg);
        break;
      case FX_PUSH_LOAD_TMP:
        L.slot_put(acc, sp++);
      case FX_LOAD_TMP:
        L.tmp_get(acc, ins.arg);
        break;
      case FX_ADD_TMP:
        L.tmp_get(b, ins.arg);
#pragma unroll
        for (int k = 0; k < K; k++) acc[k] = fadd(acc[k], b[k], F);
        break;
      case FX_MUL_TMP:
        L.tmp_get(b, ins.arg);
#pragma unroll
        for (int k = 0; k < K; k++) acc[k] = fmul(acc[k], b[k], F);
        break;
      case FX_STORE:
        L.store_row(ins.arg, acc);
        break;

      // ---- fused runs (compiler.cpp, fuse()) -----------------------------------------------------------
      case FX_TERM: {
        // T = var^e [* c]; ACC = T | ACC + T | ACC - T | T - ACC | ACC * T.  The old ACC waits in the next free
        // spill slot while the power runs (the unfused sequence pushed it there too).
        const u32 mode = (ins.flags >> TM_SHIFT) & 7u;
        if ((ins.arg >> 16) == 1u && mode != TM_LOAD) {
          // c * x as a term: nothing to park
          L.load_var(ins.arg & 0xffffu, b);
          if (ins.flags & F_CMUL) {
            const R c = mkR(ins.imm);
#pragma unroll
            for (int k = 0; k < K; k++) b[k] = fmul(b[k], c, F);
          }
          if (mode == TM_ADD) {
#pragma unroll
            for (int k = 0; k < K; k++) acc[k] = fadd(acc[k], b[k], F);
          } else if (mode == TM_SUB) {
#pragma unroll
            for (int k = 0; k < K; k++) acc[k] = fsub(acc[k], b[k], F);
          } else if (mode == TM_RSUB) {
#pragma unroll
            for (int k = 0; k < K; k++) acc[k] = fsub(b[k], acc[k], F);
          } else {
#pragma unroll
            for (int k = 0; k < K; k++) acc[k] = fmul(acc[k], b[k], F);
          }
          break;
        }
        if (mode != TM_LOAD) L.slot_put(acc, sp);
        else if (ins.flags & F_PUSH) L.slot_put(acc, sp++);
        if ((ins.arg >> 16) == 1u) {
          // ACC = c * x: the row goes straight into the accumulator (mode is TM_LOAD here)
          L.load_var(ins.arg & 0xffffu, acc);
          if (ins.flags & F_CMUL) {
            const R c = mkR(ins.imm);
#pragma unroll
            for (int k = 0; k < K; k++) acc[k] = fmul(acc[k], c, F);
          }
          break;
        }
        L.load_var(ins.arg & 0xffffu, b);
        pow_from<K>(acc, b, ins.arg >> 16, F);
        if (ins.flags & F_CMUL) {
          const R c = mkR(ins.imm);
#pragma unroll
          for (int k = 0; k < K; k++) acc[k] = fmul(acc[k], c, F);
        }
        if (mode != TM_LOAD) {
          L.slot_get(b, sp);
          if (mode == TM_ADD) {
#pragma unroll
            for (int k = 0; k < K; k++) acc[k] = fadd(b[k], acc[k], F);
          } else if (mode == TM_SUB) {
#pragma unroll
            for (int k = 0; k < K; k++) acc[k] = fsub(b[k], acc[k], F);
          } else if (mode == TM_RSUB) {
#pragma unroll
            for (int k = 0; k < K; k++) acc[k] = fsub(acc[k], b[k], F);
          } else {
#pragma unroll
            for (int k = 0; k < K; k++) acc[k] = fmul(acc[k], b[k], F);
          }
        }
        break;
      }
      case FX_ACCUM2: {
        // every branch is uniform and sits outside the per-probe loops
        const u32 sa = ins.arg & 0xffu, sb = (ins.arg >> 8) & 0xffu;
        if (ins.flags & F_INIT) {
          if (ins.flags & F_SWAP) {
#pragma unroll
            for (int k = 0; k < K; k++) b[k] = fneg(acc[k], F);
            L.tmp_put(b, sa);
          } else {
            L.tmp_put(acc, sa);
          }
        } else {
          L.tmp_get(b, sa);
          if (ins.flags & F_SWAP) {
#pragma unroll
            for (int k = 0; k < K; k++) b[k] = fsub(b[k], acc[k], F);
          } else {
#pragma unroll
            for (int k = 0; k < K; k++) b[k] = fadd(b[k], acc[k], F);
          }
          L.tmp_put(b, sa);
        }
        if (ins.flags & F_INIT2) {
          if (ins.flags & F_SWAP2) {
#pragma unroll
            for (int k = 0; k < K; k++) b[k] = fneg(acc[k], F);
            L.tmp_put(b, sb);
          } else {
            L.tmp_put(acc, sb);
          }
        } else {
          L.tmp_get(b, sb);
          if (ins.flags & F_SWAP2) {
#pragma unroll
            for (int k = 0; k < K; k++) b[k] = fsub(b[k], acc[k], F);
          } else {
#pragma unroll
            for (int k = 0; k < K; k++) b[k] = fadd(b[k], acc[k], F);
          }
          L.tmp_put(b, sb);
        }
        break;
      }
      case FX_ADDVARS: {
        const u32 cnt = ins.arg & 0xffu;
        u32 signs = ins.arg >> 8;
        u64 ids = ins.imm;
#pragma unroll 1
        for (u32 i = 0; i < cnt; i++, ids >>= 8, signs >>= 1) {
          L.load_var((u32)ids & 0xffu, b);
          if (signs & 1u) {
#pragma unroll
            for (int k = 0; k < K; k++) acc[k] = fsub(acc[k], b[k], F);
          } else {
#pragma unroll
            for (int k = 0; k < K; k++) acc[k] = fadd(acc[k], b[k], F);
          }
        }
        break;
      }
      case FX_MULVARS: {
        const u32 cnt = ins.arg & 0xffu;
        u64 ids = ins.imm;
#pragma unroll 1
        for (u32 i = 0; i < cnt; i++, ids >>= 8) {
          L.load_var((u32)ids & 0xffu, b);
#pragma unroll
          for (int k = 0; k < K; k++) acc[k] = fmul(acc[k], b[k], F);
        }
        break;
      }

      case FX_MULADD: {
        // Horner step: ACC = ACC * var + c  (F_SWAP: c - ACC * var)
        L.load_var(ins.arg, b);
        const R c = mkR(ins.imm);
#pragma unroll
        for (int k = 0; k < K; k++) acc[k] = fmul(acc[k], b[k], F);
        if (ins.flags & F_SWAP) {
#pragma unroll
          for (int k = 0; k < K; k++) acc[k] = fsub(c, acc[k], F);
        } else {
#pragma unroll
          for (int k = 0; k < K; k++) acc[k] = fadd(acc[k], c, F);
        }
        break;
      }

      // ---- everything else, out of line (copies of the state: the originals never have their address taken)
      default: {
        if (ins.op == OP_HALT) return;
        if (LEAN) {
          asm volatile("trap;");       // the launcher only picks a lean kernel for tapes without generic instructions
          break;
        }
        Lane<K, THREADS, VEC, FT> L2 = L;
        R t[K];
#pragma unroll
        for (int k = 0; k < K; k++) t[k] = acc[k];
        sp = generic_op<K, THREADS, VEC, FT>(L2, cur.v.x, cur.v.y, ins.imm, t, sp);
#pragma unroll
        for (int k = 0; k < K; k++) acc[k] = opaque(toU(t[k]));
        break;
      }
    }
    // The fetched instruction is taken over HERE, at the end of the iteration.  Left to itself ptxas copies it right
    // behind the load at the top (the fields of the current instruction have been extracted by then) and every
    // iteration waits out the whole fetch: 40 % of the stall samples of a small launch, where one warp walks the tape
    // with nothing else to hide it (s_y_4_v at 1..128 points 199 -> 156 us, f4 12.1 -> 10.3 us; +2.4 % on the
    // 2000-function collection at 65 536 points; profiles/r02_kbench_late_copy_*.log).
    asm volatile("mov.b32 %0, %4;\n\tmov.b32 %1, %5;\n\tmov.b32 %2, %6;\n\tmov.b32 %3, %7;"
                 : "=r"(cur.v.x), "=r"(cur.v.y), "=r"(cur.v.z), "=r"(cur.v.w)
                 : "r"(nxt.v.x), "r"(nxt.v.y), "r"(nxt.v.z), "r"(nxt.v.w));
    ++pc;
  }
}

template <int K, int THREADS, bool VEC, bool LEAN, class FT = Field>
cudaError_t launch_k(const EvalParams& prm, const LaunchShape& sh, cudaStream_t stream) {
  auto kern = eval_tape_kernel<K, THREADS, VEC, LEAN, FT>;
  if (sh.smem > 48 * 1024) {
    const cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sh.smem);
    if (e != cudaSuccess) return e;
  }
  kern<<<dim3(sh.blocks, sh.chunks), THREADS, sh.smem, stream>>>(prm);
  return cudaGetLastError();
}

template <int K, int THREADS>
cudaError_t launch_t(const EvalParams& prm, const LaunchShape& sh, cudaStream_t stream) {
  const bool vec = K >= 2 && (prm.n_points % 2 == 0) && (prm.in_stride % 2 == 0) && (prm.out_stride % 2 == 0) &&
                   ((uintptr_t)prm.values % 16 == 0) && ((uintptr_t)prm.out % 16 == 0) &&
                   (!prm.table || (prm.table_stride % 2 == 0 && (uintptr_t)prm.table % 16 == 0));
  constexpr bool HAS_LEAN = K == 8;                        // the only shapes whose register budget differs
  constexpr bool HAS_FULL = !(K == 8 && THREADS > 768);    // 1024 threads leave 64 registers: lean tapes only
  if (prm.generic) {
    // any odd prime: the shapes small and medium batches run in (choose_shape keeps to them)
    constexpr bool HAS_GENERIC = K <= 4 && THREADS <= 256;
    if (!HAS_GENERIC) return cudaErrorInvalidValue;
    if (vec) return launch_k<(HAS_GENERIC ? K : 1), (HAS_GENERIC ? THREADS : 128), (K >= 2 && HAS_GENERIC), false, FieldB>(prm, sh, stream);
    return launch_k<(HAS_GENERIC ? K : 1), (HAS_GENERIC ? THREADS : 128), false, false, FieldB>(prm, sh, stream);
  }
  if (sh.lean && HAS_LEAN) {
    if (vec) return launch_k<K, THREADS, (K >= 2), HAS_LEAN>(prm, sh, stream);
    return launch_k<K, THREADS, false, HAS_LEAN>(prm, sh, stream);
  }
  if (!HAS_FULL) return cudaErrorInvalidValue;
  if (vec) return launch_k<(HAS_FULL ? K : 1), THREADS, (K >= 2 && HAS_FULL), false>(prm, sh, stream);
  return launch_k<(HAS_FULL ? K : 1), THREADS, false, false>(prm, sh, stream);
}

}  // namespace

LaunchShape choose_shape(uint64_t n_points, uint32_t n_slots, int sm_count, int force_k, int force_threads,
                         unsigned max_chunks, bool lean_ok, bool generic_field) {
  LaunchShape s;
  if (generic_field) {
    // the generic-modulus kernels exist for 1, 2 and 4 probes per thread and blocks of 128 or 256 threads
    lean_ok = false;
    if (force_k > 4) force_k = 4;
    if (force_threads > 256) force_threads = 256;
  }
  auto smem_of = [&](int k, int th) { return (size_t)(64 + (size_t)n_slots * k * th) * sizeof(uint64_t); };
  // resident warps per SM for a shape: registers (launch bounds: 80 per thread at 8 probes per thread, 64
  // below) and shared memory (227 KB per SM, 1 KB reserved per block)
  auto blocks_by_smem = [&](int k, int th) { return (int)((227 * 1024) / (smem_of(k, th) + 1024)); };
  // a lean kernel (64 registers at k = 8) pays off only if shared memory lets the fourth block in
  auto lean_of = [&](int k, int th) { return lean_ok && k == 8 && blocks_by_smem(k, th) >= 1024 / th; };
  auto warps_of = [&](int k, int th) {
    const int by_regs = (k == 8 && !lean_of(k, th) ? 768 : 1024) / th;
    return std::min(by_regs, blocks_by_smem(k, th)) * th / 32;
  };
  if (force_k || force_threads) {
    s.threads = force_threads ? force_threads : 256;
    s.k = force_k ? force_k : 8;
    // the slots must fit in shared memory: back off K, then the block size
    while (smem_of(s.k, s.threads) > 200 * 1024 && s.k > 1) s.k >>= 1;
    while (smem_of(s.k, s.threads) > 200 * 1024 && s.threads > 128) s.threads >>= 1;
  } else {
    // small batches: spread over all SMs first (one point per lane), then add work per thread
    const uint64_t lanes = (uint64_t)sm_count * 1024;      // 4 CTAs of 256 threads per SM
    // (a many-function collection also spreads over grid.y, one run of functions per block)
    const uint64_t par = n_points * (uint64_t)(max_chunks ? max_chunks : 1);
    int kmax = par >= 8 * lanes ? 8 : par >= 4 * lanes ? 4 : par >= 2 * lanes ? 2 : 1;
    if (generic_field && kmax > 4) kmax = 4;
    while (kmax > 1 && (uint64_t)kmax * 128 > n_points + 127) kmax >>= 1;   // never more lanes than points
    // most probes per thread first, but never below 16 resident warps per SM if a smaller shape gets there
    s.k = 1;
    s.threads = 128;
    int best = -1;
    for (int k = kmax; k >= 1 && best < 0; k >>= 1)
      for (int th : {256, 128})
        if (smem_of(k, th) <= 200 * 1024 && warps_of(k, th) >= (k == 8 ? 16 : 24)) {
          s.k = k;
          s.threads = th;
          best = 1;
          break;
        }
    if (best < 0) {
      // nothing reaches the target: take the shape with the most resident probes
      long score = -1;
      for (int k = kmax; k >= 1; k >>= 1)
        for (int th : {256, 128}) {
          if (smem_of(k, th) > 200 * 1024) continue;
          const long sc = (long)warps_of(k, th) * k;
          if (sc > score) { score = sc; s.k = k; s.threads = th; }
        }
    }
  }
  s.smem = smem_of(s.k, s.threads);
  s.lean = lean_of(s.k, s.threads);
  const uint64_t per_block = (uint64_t)s.threads * s.k;
  s.blocks = (unsigned)((n_points + per_block - 1) / per_block);
  // too few point tiles to fill the chip: split the functions of the collection over grid.y
  s.chunks = 1;
  const unsigned want = (unsigned)sm_count * 4;
  while (s.chunks < max_chunks && s.blocks * s.chunks < want) s.chunks <<= 1;
  if (s.chunks > max_chunks) s.chunks = max_chunks;
  return s;
}

cudaError_t launch_eval(const EvalParams& prm, const LaunchShape& sh, cudaStream_t stream) {
  if (prm.n_points == 0) return cudaSuccess;
  if (sh.smem > 227 * 1024) return cudaErrorInvalidValue;
#define FFGPU_CASE(KK, TT) \
  if (sh.k == KK && sh.threads == TT) return launch_t<KK, TT>(prm, sh, stream);
  FFGPU_CASE(1, 128) FFGPU_CASE(1, 256) FFGPU_CASE(1, 512)
  FFGPU_CASE(2, 128) FFGPU_CASE(2, 256) FFGPU_CASE(2, 512)
  FFGPU_CASE(4, 128) FFGPU_CASE(4, 256) FFGPU_CASE(4, 512)
  FFGPU_CASE(8, 128) FFGPU_CASE(8, 256) FFGPU_CASE(8, 512)
  FFGPU_CASE(1, 1024) FFGPU_CASE(2, 1024) FFGPU_CASE(4, 1024) FFGPU_CASE(8, 1024)
#undef FFGPU_CASE
  return cudaErrorInvalidValue;
}

// ---- synthetic points ----------------------------------------------------------------------
namespace {
__host__ __device__ inline u64 splitmix64_at(u64 x) {
  u64 z = x + 0x9E3779B97F4A7C15ULL;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
  return z ^ (z >> 31);
}
__global__ void fill_points_kernel(u64* values, u64 n_vars, u64 n_points, u64 stride, u64 first, u64 total, u64 seed,
                                   u64 p) {
  const u64 n = n_vars * n_points;
  for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (u64)gridDim.x * blockDim.x) {
    const u64 v = i / n_points, j = i - v * n_points;
    values[v * stride + j] = 1 + splitmix64_at(seed + v * total + first + j) % (p - 1);
  }
}
__device__ __forceinline__ u64 reduce_any(u64 v, const Field& F) { return ff63::canon_pm(v, F.c, F.p); }
__device__ __forceinline__ u64 reduce_any(u64 v, const FieldB& F) { return v >= F.p ? v % F.p : v; }
// probe points on a line: values[v][j] = a[v] * t[j] + s[v]  (canonical).  What Reconstructor::queue_probes
// computes on the host per point (reference: Reconstructor.hpp:2946-2950 -- values[0] = t + shift[0],
// values[i] = rand_zi[i-1] * t + shift[i]; the factor-scan variant :2952-2960 is a[v] in {0,1}).
// Two points per thread, 16-byte accesses; a[] and s[] are read through the uniform path.
template <class FT>
__global__ void __launch_bounds__(256) line_points_kernel(u64* __restrict__ values, const u64* __restrict__ t,
                                                          const u64* __restrict__ a, const u64* __restrict__ s,
                                                          u32 n_vars, u64 n_points, u64 stride, const FT F) {
  const bool vec = (n_points % 2 == 0) && (stride % 2 == 0) && ((uintptr_t)values % 16 == 0) && ((uintptr_t)t % 16 == 0);
  if (vec) {
    for (u64 j = 2 * ((u64)blockIdx.x * blockDim.x + threadIdx.x); j < n_points; j += 2 * (u64)gridDim.x * blockDim.x) {
      ulonglong2 tt = __ldg(reinterpret_cast<const ulonglong2*>(t + j));
      tt.x = reduce_any(tt.x, F);                   // any 64-bit t is reduced first: the product needs bounded operands
      tt.y = reduce_any(tt.y, F);
      for (u32 v = 0; v < n_vars; v++) {
        const u64 av = __ldg(a + v), sv = __ldg(s + v);
        ulonglong2 x;
        x.x = fcanon(fadd(fmul(mkR(tt.x), mkR(av), F), mkR(sv), F), F);
        x.y = fcanon(fadd(fmul(mkR(tt.y), mkR(av), F), mkR(sv), F), F);
        *reinterpret_cast<ulonglong2*>(values + (u64)v * stride + j) = x;
      }
    }
  } else {
    for (u64 j = (u64)blockIdx.x * blockDim.x + threadIdx.x; j < n_points; j += (u64)gridDim.x * blockDim.x) {
      const u64 tj = reduce_any(__ldg(t + j), F);
      for (u32 v = 0; v < n_vars; v++)
        values[(u64)v * stride + j] = fcanon(fadd(fmul(mkR(tj), mkR(__ldg(a + v)), F), mkR(__ldg(s + v)), F), F);
    }
  }
}
__global__ void checksum_kernel(const u64* out, u64 n_points, u64 stride, u64 p, u64* sums) {
  // one block per output row; sum mod p (p < 2^63: a + b cannot wrap) and xor
  const u64* row = out + (u64)blockIdx.x * stride;
  u64 s = 0, x = 0;
  for (u64 j = threadIdx.x; j < n_points; j += blockDim.x) {
    const u64 v = row[j];
    s += v;
    if (s >= p) s -= p;
    x ^= v;
  }
  __shared__ u64 ss[256], xs[256];
  ss[threadIdx.x] = s;
  xs[threadIdx.x] = x;
  __syncthreads();
  for (unsigned h = blockDim.x / 2; h; h >>= 1) {
    if (threadIdx.x < h) {
      u64 a = ss[threadIdx.x] + ss[threadIdx.x + h];
      if (a >= p) a -= p;
      ss[threadIdx.x] = a;
      xs[threadIdx.x] ^= xs[threadIdx.x + h];
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    sums[2 * blockIdx.x] = ss[0];
    sums[2 * blockIdx.x + 1] = xs[0];
  }
}
template <class FT>
__global__ void field_ops_kernel(int op, const u64* a, const u64* b, u64* r, u64 n, const FT F) {
  const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
  // all threads stay alive: op 6 uses block-wide barriers
  const R x = fin(mkR(i < n ? a[i] : 1), F), y = fin(mkR(i < n ? b[i] : 1), F);
  __shared__ u64 scratch[64];
  u64 v = 0;
  switch (op) {
    case 0: v = fcanon(fadd(x, y, F), F); break;
    case 1: v = fcanon(fsub(x, y, F), F); break;
    case 2: v = fcanon(fmul(x, y, F), F); break;
    case 3: v = fcanon(fneg(x, F), F); break;
    case 4: v = fcanon(inv_fermat(x, F), F); break;
    case 5: v = fcanon(pow_ff_lane(x, y, F), F); break;
    case 6: { R t[1] = {x}; block_inverse<1, 256>(t, F, scratch); v = fcanon(t[0], F); break; }
    case 7: v = fcanon(pow_negint_lane(x, y, F), F); break;
    case 8: v = fcanon(x, F); break;
    case 9: { R t[2] = {x, y}; block_inverse<2, 256>(t, F, scratch); v = fcanon(fmul(t[0], t[1], F), F); break; }
    // chains: the result of one operation is the (un-canonicalised) input of the next
    case 10: v = fcanon(fmul(fadd(x, y, F), y, F), F); break;
    case 11: v = fcanon(fmul(fsub(x, y, F), y, F), F); break;
    case 12: { const R m = fmul(x, y, F); v = fcanon(fsub(fsqr(m, F), y, F), F); break; }
    case 13: v = fcanon(fadd(fmul(fneg(x, F), y, F), x, F), F); break;
    default: break;
  }
  if (i < n) r[i] = v;
}
}  // namespace

bool is_table_form(uint64_t p) {
  const uint64_t two63 = 1ULL << 63;
  return p < two63 && two63 - p < (1u << 15);
}
static Field host_field(uint64_t p) {
  const u32 c = (u32)((1ULL << 63) - p);
  return Field{p, c, 2 * c};
}
static FieldB host_field_b(uint64_t p) {
  const ff63::Barrett b = ff63::make_barrett(p);
  return FieldB{b.p, b.pn, b.pinv, b.k};
}
void barrett_constants(uint64_t p, uint64_t* pinv, uint32_t* shift) {
  const ff63::Barrett b = ff63::make_barrett(p);
  *pinv = b.pinv;
  *shift = b.k;
}

// SMs of the calling thread's current device (grid sizes of the grid-stride kernels)
static unsigned current_sm_count() {
  int dev = 0, n = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
    return 148;
  return (unsigned)n;
}

cudaError_t launch_fill_points(uint64_t* values, uint64_t n_vars, uint64_t n_points, uint64_t stride,
                               uint64_t first_point, uint64_t total_points, uint64_t seed, uint64_t p,
                               cudaStream_t stream) {
  if (n_vars * n_points == 0) return cudaSuccess;
  fill_points_kernel<<<current_sm_count() * 8, 256, 0, stream>>>(reinterpret_cast<u64*>(values), n_vars, n_points, stride, first_point, total_points, seed, p);
  return cudaGetLastError();
}

cudaError_t launch_line_points(uint64_t* values, const uint64_t* t, const uint64_t* a, const uint64_t* s,
                               uint32_t n_vars, uint64_t n_points, uint64_t stride, uint64_t p, cudaStream_t stream) {
  if ((uint64_t)n_vars * n_points == 0) return cudaSuccess;
  const uint64_t want = (n_points / 2 + 255) / 256;
  const unsigned blocks = (unsigned)std::min<uint64_t>(std::max<uint64_t>(want, 1), (uint64_t)current_sm_count() * 16);
  if (is_table_form(p))
    line_points_kernel<Field><<<blocks, 256, 0, stream>>>(reinterpret_cast<u64*>(values), reinterpret_cast<const u64*>(t),
                                                          reinterpret_cast<const u64*>(a), reinterpret_cast<const u64*>(s),
                                                          n_vars, n_points, stride, host_field(p));
  else
    line_points_kernel<FieldB><<<blocks, 256, 0, stream>>>(reinterpret_cast<u64*>(values), reinterpret_cast<const u64*>(t),
                                                           reinterpret_cast<const u64*>(a), reinterpret_cast<const u64*>(s),
                                                           n_vars, n_points, stride, host_field_b(p));
  return cudaGetLastError();
}

cudaError_t launch_checksum(const uint64_t* out, uint64_t n_out, uint64_t n_points, uint64_t stride, uint64_t p,
                            uint64_t* sums, cudaStream_t stream) {
  if (n_out == 0) return cudaSuccess;
  checksum_kernel<<<(unsigned)n_out, 256, 0, stream>>>(reinterpret_cast<const u64*>(out), n_points, stride, p,
                                                        reinterpret_cast<u64*>(sums));
  return cudaGetLastError();
}

cudaError_t launch_field_ops(int op, const uint64_t* a, const uint64_t* b, uint64_t* r, uint64_t n, uint64_t p,
                             cudaStream_t stream) {
  if (n == 0) return cudaSuccess;
  const unsigned blocks = (unsigned)((n + 255) / 256);
  if (is_table_form(p))
    field_ops_kernel<Field><<<blocks, 256, 0, stream>>>(op, reinterpret_cast<const u64*>(a), reinterpret_cast<const u64*>(b),
                                                        reinterpret_cast<u64*>(r), n, host_field(p));
  else
    field_ops_kernel<FieldB><<<blocks, 256, 0, stream>>>(op, reinterpret_cast<const u64*>(a), reinterpret_cast<const u64*>(b),
                                                         reinterpret_cast<u64*>(r), n, host_field_b(p));
  return cudaGetLastError();
}

}  // namespace ffgpu
