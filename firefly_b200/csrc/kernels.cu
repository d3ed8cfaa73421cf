// kernels.cu -- the sm_100a tape interpreter: batched black-box probe evaluation mod p.
//
// Replaces the reference's hot loop, ShuntingYardParser::evaluate_pre<FFInt / FFIntVec<N>>
// (reference: source/include/firefly/ShuntingYardParser.hpp:365-479) together with the FFInt
// operators it calls (source/FFInt.cpp:130-192) for ALL probe points of a batch at once.
//
// Mapping to the machine
//   * one probe point per lane, K points per thread (independent multiply chains for ILP);
//   * the accumulator of the tape machine (tape.hpp) is K 64-bit registers; intermediate
//     values are LAZY residues (any representative below 2^64, see ff63.cuh) and are made
//     canonical only where the reference's semantics need it: zero tests of divisors, lane
//     exponents, and stores;
//   * the tape is warp-uniform: every branch of the interpreter is uniform, instructions are
//     fetched with one 128-bit load one instruction ahead of execution;
//   * the spill stack lives in shared memory, laid out [slot][k][thread] so that a warp's
//     access is one conflict-free 256-byte row;
//   * probe coordinates are read as 16-byte vectors, [var][point] layout, fully coalesced;
//     results are written [function][point], 16-byte vectors, fully coalesced;
//   * divisions: ONE Fermat inversion per thread block.  Every lane's divisor enters a
//     three-level product tree (K values per thread -> 32 lanes by shuffle scans -> the
//     block's warps through shared memory); lane w of warp 0 inverts the running product of
//     warps 0..w, so the 32 lanes of the inverting warp all do useful, different work.
//     Zero divisors are replaced by 1 in the tree and produce 0 (reference: a/0 = 0).
#include "kernels.cuh"

#include <algorithm>

#include "ff63.cuh"

namespace ffgpu {

using ff63::u32;
using ff63::u64;

namespace {

struct Field {
  u64 p;
  u32 c, d;
};

__device__ __forceinline__ u64 fmul(u64 a, u64 b, const Field& F) { return ff63::mul_pm_dev(a, b, F.d); }
__device__ __forceinline__ u64 fsqr(u64 a, const Field& F) { return ff63::sqr_pm_dev(a, F.d); }
__device__ __forceinline__ u64 fadd(u64 a, u64 b, const Field& F) { return ff63::add_lazy_dev(a, b, F.d); }
__device__ __forceinline__ u64 fsub(u64 a, u64 b, const Field& F) { return ff63::sub_lazy_dev(a, b, F.d, F.p); }
__device__ __forceinline__ u64 fneg(u64 a, const Field& F) { return ff63::neg_lazy_dev(a, F.d, F.p); }
__device__ __forceinline__ void fmul_ip(u64& a, u64 b, const Field& F) { ff63::mul_pm_ip(a, b, F.d); }
__device__ __forceinline__ void fsqr_ip(u64& a, const Field& F) { ff63::sqr_pm_ip(a, F.d); }
__device__ __forceinline__ u64 fcanon(u64 a, const Field& F) { return ff63::canon_pm(a, F.c, F.p); }

// x^e, e uniform across the warp, left-to-right binary; x^0 = 1 (reference mod_pow, FFInt.cpp:371-383)
template <int K>
__device__ __forceinline__ void pow_uniform(u64 (&a)[K], u64 e, const Field& F) {
  if (e == 0) {
#pragma unroll
    for (int k = 0; k < K; k++) a[k] = 1;
    return;
  }
  u64 base[K];
#pragma unroll
  for (int k = 0; k < K; k++) base[k] = a[k];
  int bit = 62 - __clzll((long long)e);          // index below the leading one
#pragma unroll 1
  for (; bit >= 0; --bit) {
#pragma unroll
    for (int k = 0; k < K; k++) fsqr_ip(a[k], F);
    if ((e >> bit) & 1) {
#pragma unroll
      for (int k = 0; k < K; k++) fmul_ip(a[k], base[k], F);
    }
  }
}

__device__ __forceinline__ u64 sqr_n(u64 x, int n, const Field& F) {
#pragma unroll 1
  for (int i = 0; i < n; i++) x = fsqr(x, F);
  return x;
}

// x^(p-2) for p = 2^63 - c:  p - 2 = (2^47 - 1) * 2^16 + (2^16 - c - 2).
// 2^47 - 1 by the addition chain 1,2,3,5,10,20,23,46,47 on the run length; then 16 squarings
// interleaved with the bits of the 16-bit tail.  62 squarings + ~16 multiplications.  0 -> 0.
__device__ __noinline__ u64 inv_fermat(u64 x, Field F) {
  const u64 x1 = x;
  const u64 x2 = fmul(fsqr(x1, F), x1, F);
  const u64 x3 = fmul(fsqr(x2, F), x1, F);
  const u64 x5 = fmul(sqr_n(x3, 2, F), x2, F);
  const u64 x10 = fmul(sqr_n(x5, 5, F), x5, F);
  const u64 x20 = fmul(sqr_n(x10, 10, F), x10, F);
  const u64 x23 = fmul(sqr_n(x20, 3, F), x3, F);
  const u64 x46 = fmul(sqr_n(x23, 23, F), x23, F);
  u64 r = fmul(fsqr(x46, F), x1, F);             // x^(2^47 - 1)
  const u32 tail = 65536u - F.c - 2u;            // c < 2^15: 16 bits
#pragma unroll 1
  for (int bit = 15; bit >= 0; --bit) {
    r = fsqr(r, F);
    if ((tail >> bit) & 1) r = fmul(r, x1, F);
  }
  return r;
}

// per-lane power with a lane-dependent exponent (divergent; only for `x^y` style input)
__device__ __noinline__ u64 pow_lane(u64 b, u64 e, Field F) {
  u64 r = 1;
#pragma unroll 1
  while (e) {
    if (e & 1) r = fmul(r, b, F);
    b = fsqr(b, F);
    e >>= 1;
  }
  return r;
}
// FFInt::pow(const FFInt&) per lane (FFInt.cpp:155-179)
__device__ __forceinline__ u64 pow_ff_lane(u64 b, u64 e_lazy, const Field& F) {
  const u64 e = fcanon(e_lazy, F);
  if (e == 2) return fsqr(b, F);
  if (e < (F.p >> 1)) return pow_lane(b, e, F);
  return pow_lane(inv_fermat(b, F), F.p - e - 1, F);
}
// FFInt::pow<int>(to_neg_int()) per lane (FFInt.hpp:149-174, FFInt.cpp:226-228)
__device__ __forceinline__ u64 pow_negint_lane(u64 b, u64 e_lazy, const Field& F) {
  const int e = (int)(0u - (u32)fcanon(e_lazy, F));
  if (e == 2) return fsqr(b, F);
  if (e >= 0) return pow_lane(b, (u64)e, F);
  return pow_lane(inv_fermat(b, F), (u64)(-(long long)e), F);
}

// ---- block-batched inversion -----------------------------------------------------------------
// a[k] <- 1/a[k] (0 -> 0) for the K values of every thread of the block.  scratch: 64 u64 of
// shared memory.  Must be reached by all threads of the block (the tape is block uniform).
template <int K, int THREADS>
__device__ __forceinline__ void block_inverse(u64 (&a)[K], const Field& F, u64* scratch) {
  constexpr int W = THREADS / 32;
  const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const unsigned FULL = 0xffffffffu;
  bool z[K];
  u64 e[K], pre[K];
#pragma unroll
  for (int k = 0; k < K; k++) {
    z[k] = fcanon(a[k], F) == 0;
    e[k] = z[k] ? 1 : a[k];
  }
  pre[0] = e[0];
#pragma unroll
  for (int k = 1; k < K; k++) pre[k] = fmul(pre[k - 1], e[k], F);
  const u64 t = pre[K - 1];
  // inclusive prefix and inclusive suffix products over the lanes of the warp
  u64 s = t, u = t;
#pragma unroll
  for (int off = 1; off < 32; off <<= 1) {
    const u64 so = __shfl_up_sync(FULL, s, off);
    const u64 uo = __shfl_down_sync(FULL, u, off);
    const u64 sm = fmul(s, so, F);
    const u64 um = fmul(u, uo, F);
    if (lane >= (unsigned)off) s = sm;
    if (lane + off < 32) u = um;
  }
  u64 sx = __shfl_up_sync(FULL, s, 1);      // product of the lanes before this one
  u64 ux = __shfl_down_sync(FULL, u, 1);    // product of the lanes after this one
  if (lane == 0) sx = 1;
  if (lane == 31) ux = 1;
  u64 inv_upto, before;                     // 1 / (warps 0..warp), product of warps 0..warp-1
  if (W > 1) {
    if (lane == 31) scratch[warp] = s;
    __syncthreads();
    if (warp == 0) {
      u64 S = lane < W ? scratch[lane] : 1;
#pragma unroll
      for (int off = 1; off < W; off <<= 1) {
        const u64 so = __shfl_up_sync(FULL, S, off);
        const u64 sm = fmul(S, so, F);
        if (lane >= (unsigned)off) S = sm;
      }
      const u64 IS = inv_fermat(S, F);
      u64 Sx = __shfl_up_sync(FULL, S, 1);
      if (lane == 0) Sx = 1;
      scratch[lane] = Sx;
      scratch[32 + lane] = IS;
    }
    __syncthreads();
    before = scratch[warp];
    inv_upto = scratch[32 + warp];
  } else {
    const u64 total = __shfl_sync(FULL, s, 31);
    inv_upto = inv_fermat(total, F);
    before = 1;
  }
  // 1/t = 1/(everything up to the end of this warp) * (warps before) * (lanes before) * (lanes after)
  u64 J = fmul(fmul(inv_upto, ux, F), fmul(before, sx, F), F);
#pragma unroll
  for (int k = K - 1; k >= 1; --k) {
    a[k] = z[k] ? 0 : fmul(J, pre[k - 1], F);
    J = fmul(J, e[k], F);
  }
  a[0] = z[0] ? 0 : J;
}

// ---- the interpreter -------------------------------------------------------------------------
union InstrWord {
  uint4 v;
  Instr i;
};

template <int K, int THREADS, bool VEC>
__global__ void __launch_bounds__(THREADS, (K == 8 ? 768 : 1024) / THREADS) eval_tape_kernel(const EvalParams prm) {
  extern __shared__ __align__(16) u64 smem[];
  u64* const scratch = smem;                 // 64 words for block_inverse
  u64* const stk = smem + 64 + threadIdx.x;  // [slot][k][thread], this thread's column
  const Field F{prm.p, prm.c, prm.d};

  // Points owned by this thread.  VEC: pairs (2t, 2t+1) inside rows of 2*THREADS points, so that
  // every load/store is one 16-byte access per lane and a warp covers 512 contiguous bytes.
  const u64 tile = (u64)blockIdx.x * (u64)(THREADS * K);
  constexpr u32 STEP = VEC ? 2 * THREADS : THREADS;       // distance between this thread's accesses
  constexpr int NACC = VEC ? K / 2 : K;                   // accesses per variable
  // Tile-relative offsets of this thread's accesses, clamped into the batch: lanes beyond the end
  // recompute the last valid point and never store, so the hot loop has no bounds checks.
  const u64 left64 = prm.n_points - tile;                 // > 0: the grid covers exactly the batch
  const u32 left = left64 < (u64)(THREADS * K) ? (u32)left64 : (u32)(THREADS * K);
  const u32 own = VEC ? 2 * threadIdx.x : threadIdx.x;
  u32 off[NACC];
#pragma unroll
  for (int j = 0; j < NACC; j++) {
    const u32 o = own + j * STEP, last = left - (VEC ? 2 : 1);
    off[j] = o < last ? o : last;
  }
  const u64* const vbase = reinterpret_cast<const u64*>(prm.values) + tile;
  u64* const obase = reinterpret_cast<u64*>(prm.out) + tile;

  auto load_var = [&](u32 var, u64 (&b)[K]) {
    const u64* row = vbase + (u64)var * prm.in_stride;
    if (VEC) {
#pragma unroll
      for (int j = 0; j < NACC; j++) {
        // n_points is even and rows are 16-byte aligned on this path
        const ulonglong2 v = __ldg(reinterpret_cast<const ulonglong2*>(row + off[j]));
        b[2 * j] = v.x;
        b[2 * j + 1] = v.y;
      }
    } else {
#pragma unroll
      for (int j = 0; j < NACC; j++) b[j] = __ldg(row + off[j]);
    }
  };
  auto in_range = [&](int j) { return own + j * STEP < left; };
  auto push = [&](const u64 (&a)[K], u32 sp) {
#pragma unroll
    for (int k = 0; k < K; k++) stk[(sp * K + k) * THREADS] = a[k];
  };
  auto pop = [&](u64 (&b)[K], u32 sp) {
#pragma unroll
    for (int k = 0; k < K; k++) b[k] = stk[(sp * K + k) * THREADS];
  };

  // temp slots (shared sub-expressions, co-evaluated sums) count down from the last slot
  const u32 tmp_top = prm.n_slots - 1;
  auto tmp_get = [&](u64 (&b)[K], u32 slot) {
    const u64* t = stk + (size_t)(tmp_top - slot) * (K * THREADS);
#pragma unroll
    for (int k = 0; k < K; k++) b[k] = t[k * THREADS];
  };
  auto tmp_put = [&](const u64 (&a)[K], u32 slot) {
    u64* t = stk + (size_t)(tmp_top - slot) * (K * THREADS);
#pragma unroll
    for (int k = 0; k < K; k++) t[k * THREADS] = a[k];
  };

  u64 acc[K];
#pragma unroll
  for (int k = 0; k < K; k++) acc[k] = 0;
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
    u64 b[K];
    switch (ins.fast) {
      // ---- the common (op, operand) combinations, each working on ACC in place -----------------
      case FX_PUSH_LOAD_VAR:
        push(acc, sp++);
      case FX_LOAD_VAR:
        load_var(ins.arg, acc);
        break;
      case FX_PUSH_LOAD_CONST:
        push(acc, sp++);
      case FX_LOAD_CONST:
#pragma unroll
        for (int k = 0; k < K; k++) acc[k] = ins.imm;
        break;
      case FX_ADD_VAR:
        load_var(ins.arg, b);
#pragma unroll
        for (int k = 0; k < K; k++) acc[k] = fadd(acc[k], b[k], F);
        break;
      case FX_ADD_CONST:
#pragma unroll
        for (int k = 0; k < K; k++) acc[k] = fadd(acc[k], ins.imm, F);
        break;
      case FX_ADD_POP:
        pop(b, --sp);
#pragma unroll
        for (int k = 0; k < K; k++) acc[k] = fadd(acc[k], b[k], F);
        break;
      case FX_SUB_VAR:
        load_var(ins.arg, b);
#pragma unroll
        for (int k = 0; k < K; k++) acc[k] = fsub(acc[k], b[k], F);
        break;
      case FX_RSUB_VAR:
        load_var(ins.arg, b);
#pragma unroll
        for (int k = 0; k < K; k++) acc[k] = fsub(b[k], acc[k], F);
        break;
      case FX_RSUB_CONST:
#pragma unroll
        for (int k = 0; k < K; k++) acc[k] = fsub(ins.imm, acc[k], F);
        break;
      case FX_SUB_POP:
        pop(b, --sp);
#pragma unroll
        for (int k = 0; k < K; k++) acc[k] = fsub(acc[k], b[k], F);
        break;
      case FX_RSUB_POP:
        pop(b, --sp);
#pragma unroll
        for (int k = 0; k < K; k++) acc[k] = fsub(b[k], acc[k], F);
        break;
      case FX_MUL_VAR:
        load_var(ins.arg, b);
#pragma unroll
        for (int k = 0; k < K; k++) fmul_ip(acc[k], b[k], F);
        break;
      case FX_MUL_CONST:
#pragma unroll
        for (int k = 0; k < K; k++) fmul_ip(acc[k], ins.imm, F);
        break;
      case FX_MUL_POP:
        pop(b, --sp);
#pragma unroll
        for (int k = 0; k < K; k++) fmul_ip(acc[k], b[k], F);
        break;
      case FX_SQR:
#pragma unroll
        for (int k = 0; k < K; k++) fsqr_ip(acc[k], F);
        break;
      case FX_POWU: {
        // FFInt::pow(const FFInt&) with a lane-uniform exponent residue (FFInt.cpp:155-179)
        const u64 e = ins.imm;
        if (e == 2) {
#pragma unroll
          for (int k = 0; k < K; k++) fsqr_ip(acc[k], F);
        } else if (e < (F.p >> 1)) {
          pow_uniform<K>(acc, e, F);
        } else {
          block_inverse<K, THREADS>(acc, F, scratch);
          pow_uniform<K>(acc, F.p - e - 1, F);
        }
        break;
      }
      case FX_INV:
        block_inverse<K, THREADS>(acc, F, scratch);
        break;
      case FX_SAVE:
        tmp_put(acc, ins.arg);
        break;
      case FX_ACCUM_INIT:
        tmp_put(acc, ins.arg);
        break;
      case FX_ACCUM_ADD:
        tmp_get(b, ins.arg);
#pragma unroll
        for (int k = 0; k < K; k++) b[k] = fadd(b[k], acc[k], F);
        tmp_put(b, ins.arg);
        break;
      case FX_ACCUM_SUB:
        tmp_get(b, ins.arg);
#pragma unroll
        for (int k = 0; k < K; k++) b[k] = fsub(b[k], acc[k], F);
        tmp_put(b, ins.arg);
        break;
      case FX_LOAD_TMP:
        tmp_get(acc, ins.arg);
        break;
      case FX_ADD_TMP:
        tmp_get(b, ins.arg);
#pragma unroll
        for (int k = 0; k < K; k++) acc[k] = fadd(acc[k], b[k], F);
        break;
      case FX_MUL_TMP:
        tmp_get(b, ins.arg);
#pragma unroll
        for (int k = 0; k < K; k++) fmul_ip(acc[k], b[k], F);
        break;
      case FX_STORE: {
        u64* row = obase + (u64)ins.arg * prm.out_stride;
        if (VEC) {
#pragma unroll
          for (int j = 0; j < NACC; j++)
            if (in_range(j))
              *reinterpret_cast<ulonglong2*>(row + off[j]) =
                  make_ulonglong2(fcanon(acc[2 * j], F), fcanon(acc[2 * j + 1], F));
        } else {
#pragma unroll
          for (int j = 0; j < NACC; j++)
            if (in_range(j)) row[off[j]] = fcanon(acc[j], F);
        }
        break;
      }

      // ---- everything else: general two-level decode ------------------------------------------------
      default: {
        if (ins.op == OP_HALT) return;
        switch (ins.src) {
          case SRC_POP: pop(b, --sp); break;
          case SRC_VAR: load_var(ins.arg, b); break;
          case SRC_NEGVAR:
            load_var(ins.arg, b);
#pragma unroll
            for (int k = 0; k < K; k++) b[k] = fneg(b[k], F);
            break;
          case SRC_CONST:
#pragma unroll
            for (int k = 0; k < K; k++) b[k] = ins.imm;
            break;
          case SRC_TMP: tmp_get(b, ins.arg); break;
          default:
#pragma unroll
            for (int k = 0; k < K; k++) b[k] = 0;
            break;
        }
        switch (ins.op) {
          case OP_LOAD:
            if (ins.flags & F_PUSH) push(acc, sp++);
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
            for (int k = 0; k < K; k++) fmul_ip(acc[k], b[k], F);
            break;
          case OP_MULADD: {
            u64 x[K];
            load_var(ins.arg, x);
#pragma unroll
            for (int k = 0; k < K; k++) acc[k] = fadd(fmul(acc[k], x[k], F), ins.imm, F);
            break;
          }
          case OP_SQR:
#pragma unroll
            for (int k = 0; k < K; k++) fsqr_ip(acc[k], F);
            break;
          case OP_SAVE:
            tmp_put(acc, ins.arg);
            break;
          case OP_ACCUM:
            if (ins.flags & F_INIT) {
#pragma unroll
              for (int k = 0; k < K; k++) b[k] = (ins.flags & F_SWAP) ? fneg(acc[k], F) : acc[k];
            } else {
              tmp_get(b, ins.arg);
#pragma unroll
              for (int k = 0; k < K; k++) b[k] = (ins.flags & F_SWAP) ? fsub(b[k], acc[k], F) : fadd(b[k], acc[k], F);
            }
            tmp_put(b, ins.arg);
            break;
          case OP_NEG:
#pragma unroll
            for (int k = 0; k < K; k++) acc[k] = fneg(acc[k], F);
            break;
          case OP_INV:
            block_inverse<K, THREADS>(acc, F, scratch);
            break;
          case OP_POWU: {
            const u64 e = ins.imm;
            if (e == 2) {
#pragma unroll
              for (int k = 0; k < K; k++) fsqr_ip(acc[k], F);
            } else if (e < (F.p >> 1)) {
              pow_uniform<K>(acc, e, F);
            } else {
              block_inverse<K, THREADS>(acc, F, scratch);
              pow_uniform<K>(acc, F.p - e - 1, F);
            }
            break;
          }
          case OP_NPOWU: {
            // FFInt::pow<int>(a.to_neg_int()) with a lane-uniform exponent (FFInt.hpp:149-174)
            const int e = (int)(0u - (u32)ins.imm);
            if (e == 2) {
#pragma unroll
              for (int k = 0; k < K; k++) fsqr_ip(acc[k], F);
            } else if (e >= 0) {
              pow_uniform<K>(acc, (u64)e, F);
            } else {
              block_inverse<K, THREADS>(acc, F, scratch);
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
          case OP_STORE: {
            u64* row = obase + (u64)ins.arg * prm.out_stride;
#pragma unroll
            for (int k = 0; k < K; k++) {
              const int j = VEC ? k / 2 : k;
              if (in_range(j)) row[off[j] + (VEC ? (k & 1) : 0)] = fcanon(acc[k], F);
            }
            break;
          }
          default:
            break;
        }
        if (ins.flags & F_NEGR) {
#pragma unroll
          for (int k = 0; k < K; k++) acc[k] = fneg(acc[k], F);
        }
        break;
      }
    }
    cur = nxt;
    ++pc;
  }
}

template <int K, int THREADS>
cudaError_t launch_t(const EvalParams& prm, const LaunchShape& sh, cudaStream_t stream) {
  const bool vec = K >= 2 && (prm.n_points % 2 == 0) && (prm.in_stride % 2 == 0) && (prm.out_stride % 2 == 0) &&
                   ((uintptr_t)prm.values % 16 == 0) && ((uintptr_t)prm.out % 16 == 0);
  cudaError_t e;
  if (vec) {
    auto kern = eval_tape_kernel<K, THREADS, (K >= 2)>;
    if (sh.smem > 48 * 1024) {
      e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sh.smem);
      if (e != cudaSuccess) return e;
    }
    kern<<<dim3(sh.blocks, sh.chunks), THREADS, sh.smem, stream>>>(prm);
  } else {
    auto kern = eval_tape_kernel<K, THREADS, false>;
    if (sh.smem > 48 * 1024) {
      e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sh.smem);
      if (e != cudaSuccess) return e;
    }
    kern<<<dim3(sh.blocks, sh.chunks), THREADS, sh.smem, stream>>>(prm);
  }
  return cudaGetLastError();
}

}  // namespace

LaunchShape choose_shape(uint64_t n_points, uint32_t n_slots, int sm_count, int force_k, int force_threads,
                         unsigned max_chunks) {
  LaunchShape s;
  auto smem_of = [&](int k, int th) { return (size_t)(64 + (size_t)n_slots * k * th) * sizeof(uint64_t); };
  // resident warps per SM for a shape: registers (launch bounds: 80 per thread at 8 probes per thread, 64
  // below) and shared memory (227 KB per SM, 1 KB reserved per block)
  auto warps_of = [&](int k, int th) {
    const int by_regs = (k == 8 ? 768 : 1024) / th;
    const int by_smem = (int)((227 * 1024) / (smem_of(k, th) + 1024));
    return std::min(by_regs, by_smem) * th / 32;
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
__global__ void field_ops_kernel(int op, const u64* a, const u64* b, u64* r, u64 n, u64 p) {
  const Field F{p, (u32)(0x8000000000000000ULL - p), (u32)(2 * (0x8000000000000000ULL - p))};
  const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
  // all threads stay alive: op 6 uses block-wide barriers
  const u64 x = i < n ? a[i] : 1, y = i < n ? b[i] : 1;
  __shared__ u64 scratch[64];
  u64 v = 0;
  switch (op) {
    case 0: v = fcanon(fadd(x, y, F), F); break;
    case 1: v = fcanon(fsub(x, y, F), F); break;
    case 2: v = fcanon(fmul(x, y, F), F); break;
    case 3: v = fcanon(fneg(x, F), F); break;
    case 4: v = fcanon(inv_fermat(x, F), F); break;
    case 5: v = fcanon(pow_ff_lane(x, y, F), F); break;
    case 6: { u64 t[1] = {x}; block_inverse<1, 256>(t, F, scratch); v = fcanon(t[0], F); break; }
    case 7: v = fcanon(pow_negint_lane(x, y, F), F); break;
    case 8: v = fcanon(x, F); break;
    case 9: { u64 t[2] = {x, y}; block_inverse<2, 256>(t, F, scratch); v = fcanon(fmul(t[0], t[1], F), F); break; }
    default: break;
  }
  if (i < n) r[i] = v;
}
}  // namespace

cudaError_t launch_fill_points(uint64_t* values, uint64_t n_vars, uint64_t n_points, uint64_t stride,
                               uint64_t first_point, uint64_t total_points, uint64_t seed, uint64_t p,
                               cudaStream_t stream) {
  if (n_vars * n_points == 0) return cudaSuccess;
  fill_points_kernel<<<148 * 8, 256, 0, stream>>>(reinterpret_cast<u64*>(values), n_vars, n_points, stride, first_point, total_points, seed, p);
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
  field_ops_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(op, reinterpret_cast<const u64*>(a),
                                                                     reinterpret_cast<const u64*>(b),
                                                                     reinterpret_cast<u64*>(r), n, p);
  return cudaGetLastError();
}

}  // namespace ffgpu
