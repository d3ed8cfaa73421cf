// ff63.cuh -- arithmetic in Z_p for 63-bit primes, host + device (sm_100a).
//
// Replaces the reference's FFInt operators (reference: source/FFInt.cpp:130-192
// add/sub/mul/div/pow/neg, :356-405 mod_mul/mod_pow/mod_inv; FFInt.hpp:149-174
// integer-exponent pow).  Results are the canonical residues the reference
// produces; HOW they are computed is different:
//
//  * FireFly's primes (reference: source/ReconstHelper.cpp:24-125) are all of the
//    form p = 2^63 - c with 25 <= c <= 13797.  2^63 = c and 2^64 = d = 2c (mod p), so a
//    128-bit product folds with two 32x32->64 multiplies by a 16-bit constant instead of
//    a long-double quotient estimate (reference DEFAULT) or a Barrett preinverse
//    (reference FLINT).  c < 2^15 is required.
//  * What the interpreter (kernels.cu) computes in is the "b63" family below: BOUNDED lazy
//    residues, any representative under 2^63 + 2^32, as two 32-bit words.  canon_pm()
//    brings a value back to [0, p) where the reference's semantics need it.
//  * Also here, for comparison in bench/pipe_peak.cu and for host-side use: the plain
//    canonical operations, the unbounded lazy forms the first kernel used (mul_pm,
//    mul_pm_dev, add_lazy ...: any representative under 2^64, 6 more adds per multiplication)
//    and a Barrett path for moduli that are not of the 2^63 - c form.
//
// The B200 integer path (measured, bench/pipe_peak.cu): IMAD.WIDE.U32 (32x32+64 -> 64) occupies
// the fma-heavy pipe for 4 cycles per warp and blocks one ALU issue slot, a 32-bit IMAD 2 cycles,
// IADD3 / LOP3 / SHF 2 cycles on the alu pipe.  One bounded multiplication is 6 wide + 1 narrow
// multiplies, one IMAD.X and 8 add/logic instructions placed on their pipes by hand (see the "b63" family
// below): 1239 G modular multiplications/s on the chip, squarings 1280.
#pragma once
#include <cstdint>

#if defined(__CUDACC__)
#define FF63_HD __host__ __device__ __forceinline__
#define FF63_D __device__ __forceinline__
#else
#define FF63_HD inline
#endif

namespace ff63 {

typedef unsigned long long u64;
typedef unsigned int u32;

// ----- small helpers ---------------------------------------------------------
FF63_HD u32 lo32(u64 x) { return (u32)x; }
FF63_HD u32 hi32(u64 x) { return (u32)(x >> 32); }

FF63_HD u64 umulhi(u64 a, u64 b) {
#if defined(__CUDA_ARCH__)
  return __umul64hi(a, b);
#else
  return (u64)(((unsigned __int128)a * b) >> 64);
#endif
}

// ----- canonical add / sub / neg (reference: FFInt.cpp:131-139, 190-192) ------
// a, b in [0,p).  p < 2^63 so a+b never wraps.
FF63_HD u64 add(u64 a, u64 b, u64 p) {
  u64 s = a + b;
  return s >= p ? s - p : s;
}
FF63_HD u64 sub(u64 a, u64 b, u64 p) {
  u64 s = a - b;
  return a < b ? s + p : s;
}
// -0 == 0: FFInt::operator-() builds FFInt(p - n) whose ctor reduces p to 0.
FF63_HD u64 neg(u64 a, u64 p) { return a ? p - a : 0; }

// ----- pseudo-Mersenne path: p = 2^63 - c, c < 2^15, d = 2c ---------------------
// lazy value: any u64 congruent to the residue.  2^64 = d (mod p).

// lazy -> canonical.  v = b*2^63 + l  ->  l + b*c < 2^63 + c < 2p.
FF63_HD u64 canon_pm(u64 v, u32 c, u64 p) {
  u64 l = (v & 0x7fffffffffffffffULL) + ((v >> 63) ? (u64)c : 0ULL);
  return l >= p ? l - p : l;
}

// fold a 128-bit value (hi:lo) to a lazy 64-bit one.
//   R  = lo + d*hi            < 2^64 + 2^80      (one 64x16 multiply = 2 IMAD.WIDE)
//   r  = R_lo + d*R_hi        R_hi <= d, d*R_hi <= d^2 < 2^32   (1 IMAD)
//   a carry out of that last add is worth another d; it cannot carry twice
//   because after a wrap r < d^2.
FF63_HD u64 fold_pm(u64 hi, u64 lo, u32 d) {
  u64 u = (u64)lo32(hi) * d;            // < 2^48
  u64 v = (u64)hi32(hi) * d;            // < 2^48
  u64 a = lo + u;
  u32 cy = a < u;
  u64 vs = v << 32;
  u64 b = a + vs;
  cy += b < vs;
  u32 rh = hi32(v) + cy;                // <= d
  u64 t = (u64)(rh * d);                // 32-bit product, fits
  u64 r = b + t;
  if (r < t) r += d;
  return r;
}

// lazy x lazy -> lazy
FF63_HD u64 mul_pm(u64 x, u64 y, u32 d) {
  return fold_pm(umulhi(x, y), x * y, d);
}
FF63_HD u64 sqr_pm(u64 x, u32 d) {
  return fold_pm(umulhi(x, x), x * x, d);
}

// canonical x canonical -> canonical (straightforward; used off the hot loop)
FF63_HD u64 mul_canon_pm(u64 a, u64 b, u32 c, u64 p) {
  return canon_pm(mul_pm(a, b, 2 * c), c, p);
}

// ----- generic path: any modulus 2 <= p < 2^63 ----------------------------------
// Barrett reduction with a normalised divisor.  Off the benchmarked path
// (FireFly's own prime table never needs it) but required so that a user prime
// set through FFInt::set_new_prime (reference: FFInt.cpp:350-352) still works.
//   k    = number of leading zero bits of p minus one, pn = p << k in [2^62, 2^63)
//   pinv = floor((2^126 - 1) / pn)  in [2^63, 2^64)
//   x = a*b < p^2;  xn = x << k < pn * p < 2^126;  q = hi64((xn >> 62) * pinv)
//   Q - 3 <= q <= Q = floor(xn / pn), so xn - q*pn in [0, 4 pn): at most three
//   corrections; the remainder is (xn mod pn) >> k.
struct Barrett {
  u64 p;      // modulus
  u64 pn;     // p << k
  u64 pinv;   // floor((2^126 - 1) / pn)
  u32 k;
};

inline Barrett make_barrett(u64 p) {
  Barrett f;
  f.p = p;
  u32 k = 0;
  while (((p << k) >> 62) == 0) ++k;
  f.k = k;
  f.pn = p << k;
  unsigned __int128 num = (((unsigned __int128)1) << 126) - 1;
  f.pinv = (u64)(num / f.pn);
  return f;
}

FF63_HD u64 mul_barrett(u64 a, u64 b, const Barrett& f) {
  u64 lo = a * b, hi = umulhi(a, b);
  if (f.k) { hi = (hi << f.k) | (lo >> (64 - f.k)); lo <<= f.k; }
  u64 xs = (hi << 2) | (lo >> 62);
  u64 q = umulhi(xs, f.pinv);
  u64 ql = q * f.pn, qh = umulhi(q, f.pn);
  u64 r = lo - ql;
  u64 rh = hi - qh - (lo < ql);
  for (int i = 0; i < 3; i++) {
    if (rh || r >= f.pn) { rh -= (r < f.pn); r -= f.pn; }
  }
  return r >> f.k;
}

// ----- lazy add / sub / neg: any representatives in [0, 2^64), p = 2^63 - c, d = 2c ----------
// 2^64 = 2p + d, so a wrap of the 64-bit sum is worth +d and a borrow is worth -d.
FF63_HD u64 add_lazy(u64 a, u64 b, u32 d) {
  u64 s = a + b;
  if (s < a) {              // wrapped once: s + 2^64 == s + d
    u64 t = s + d;
    if (t < s) t += d;      // wrapped again: t < d now, cannot wrap a third time
    s = t;
  }
  return s;
}
FF63_HD u64 sub_lazy(u64 a, u64 b, u32 d, u64 p) {
  u64 s = a - b;
  if (a < b) {              // borrowed: s - 2^64 == s - d
    u64 t = s - d;
    if (s < d) t += p;      // s - d in (-d, 0): lift by p (mod 2^64 arithmetic does the rest)
    s = t;
  }
  return s;
}
FF63_HD u64 neg_lazy(u64 a, u32 d, u64 p) { return sub_lazy(0, a, d, p); }

#if defined(__CUDACC__)
// ----- device-only tuned variants -------------------------------------------------
// Same values as mul_pm / sqr_pm; written in PTX so that NVVM cannot "improve"
// the limb arithmetic (it inserts known-bit masks on squares) and so that the
// multiply/add mix that reaches ptxas is the one counted in DESIGN.md.
FF63_D u64 mul_pm_dev(u64 x, u64 y, u32 d) {
  u64 r;
  asm("{\n\t"
      ".reg .u32 x0,x1,y0,y1,w0,w1,w2,w3,u0,u1,v0,v1,r0,r1,r2,t;\n\t"
      ".reg .u64 t0,ma,mb,t3,u,v;\n\t"
      ".reg .pred p;\n\t"
      "mov.b64 {x0,x1}, %1;\n\t"
      "mov.b64 {y0,y1}, %2;\n\t"
      "mul.wide.u32 t0, x0, y0;\n\t"
      "mul.wide.u32 ma, x0, y1;\n\t"
      "mul.wide.u32 mb, x1, y0;\n\t"
      "mul.wide.u32 t3, x1, y1;\n\t"
      "mov.b64 {w0,w1}, t0;\n\t"
      "mov.b64 {w2,w3}, t3;\n\t"
      "mov.b64 {u0,u1}, ma;\n\t"
      "mov.b64 {v0,v1}, mb;\n\t"
      "add.cc.u32 w1, w1, u0;\n\t"
      "addc.cc.u32 w2, w2, u1;\n\t"
      "addc.u32 w3, w3, 0;\n\t"
      "add.cc.u32 w1, w1, v0;\n\t"
      "addc.cc.u32 w2, w2, v1;\n\t"
      "addc.u32 w3, w3, 0;\n\t"
      // fold 1: (w1:w0) + d*w2 + (d*w3 << 32)
      "mul.wide.u32 u, w2, %3;\n\t"
      "mul.wide.u32 v, w3, %3;\n\t"
      "mov.b64 {u0,u1}, u;\n\t"
      "mov.b64 {v0,v1}, v;\n\t"
      "add.cc.u32 r0, w0, u0;\n\t"
      "addc.cc.u32 r1, w1, u1;\n\t"
      "addc.u32 r2, v1, 0;\n\t"
      "add.cc.u32 r1, r1, v0;\n\t"
      "addc.u32 r2, r2, 0;\n\t"
      // fold 2: + d*r2 (< 2^32), one conditional +d on wrap
      "mul.lo.u32 t, r2, %3;\n\t"
      "add.cc.u32 r0, r0, t;\n\t"
      "addc.cc.u32 r1, r1, 0;\n\t"
      "addc.u32 t, 0, 0;\n\t"
      "setp.ne.u32 p, t, 0;\n\t"
      "@p add.u32 r0, r0, %3;\n\t"
      "mov.b64 %0, {r0,r1};\n\t"
      "}" : "=l"(r) : "l"(x), "l"(y), "r"(d));
  return r;
}

// squaring: x0*x1 is computed once and added twice (5 wide multiplies + 1 narrow instead of 6 + 1)
FF63_D u64 sqr_pm_dev(u64 x, u32 d) {
  u64 r;
  asm("{\n\t"
      ".reg .u32 x0,x1,w0,w1,w2,w3,u0,u1,v0,v1,r0,r1,r2,t;\n\t"
      ".reg .u64 t0,m,t3,u,v;\n\t"
      ".reg .pred p;\n\t"
      "mov.b64 {x0,x1}, %1;\n\t"
      "mul.wide.u32 t0, x0, x0;\n\t"
      "mul.wide.u32 m, x0, x1;\n\t"
      "mul.wide.u32 t3, x1, x1;\n\t"
      "mov.b64 {w0,w1}, t0;\n\t"
      "mov.b64 {w2,w3}, t3;\n\t"
      "mov.b64 {u0,u1}, m;\n\t"
      "add.cc.u32 w1, w1, u0;\n\t"
      "addc.cc.u32 w2, w2, u1;\n\t"
      "addc.u32 w3, w3, 0;\n\t"
      "add.cc.u32 w1, w1, u0;\n\t"
      "addc.cc.u32 w2, w2, u1;\n\t"
      "addc.u32 w3, w3, 0;\n\t"
      "mul.wide.u32 u, w2, %2;\n\t"
      "mul.wide.u32 v, w3, %2;\n\t"
      "mov.b64 {u0,u1}, u;\n\t"
      "mov.b64 {v0,v1}, v;\n\t"
      "add.cc.u32 r0, w0, u0;\n\t"
      "addc.cc.u32 r1, w1, u1;\n\t"
      "addc.u32 r2, v1, 0;\n\t"
      "add.cc.u32 r1, r1, v0;\n\t"
      "addc.u32 r2, r2, 0;\n\t"
      "mul.lo.u32 t, r2, %2;\n\t"
      "add.cc.u32 r0, r0, t;\n\t"
      "addc.cc.u32 r1, r1, 0;\n\t"
      "addc.u32 t, 0, 0;\n\t"
      "setp.ne.u32 p, t, 0;\n\t"
      "@p add.u32 r0, r0, %2;\n\t"
      "mov.b64 %0, {r0,r1};\n\t"
      "}" : "=l"(r) : "l"(x), "r"(d));
  return r;
}


// ----- "b63" family: BOUNDED lazy residues, the representation the interpreter computes in -----------
// Every value v satisfies v < B = 2^63 + 2^32 (so its high word is <= 2^31); canonical residues qualify.
// Why: with both high words <= 2^31 the two cross products of a 64x64 multiplication sum to less than
// 2^64, so they are accumulated INSIDE one IMAD.WIDE (no carry chain), the top product limb stays below
// 2^30 + 1, and the second fold can be done at 2^63 = c with no conditional correction.
//
// The machine these forms are written for (bench/pipe_peak.cu, bench/modmul_forms.cu; profiles/r01_pipe_peak_v3.json,
// r01_modmul_forms_v1.json): per warp and SM sub-partition an IMAD.WIDE holds the fma-heavy pipe for 4 cycles AND
// the alu pipe for 2 (1 wide + 2 IADD3 take 6 cycles, 1 wide + 1 IMAD 6, 1 wide + 1 IADD3 4.2), +1.25 alu-side
// when it has a 64-bit addend; any other IMAD* takes 2 fma cycles, IADD3 / LOP3 / SHF / SEL 2 alu cycles.  So
//     fma = 4 W + 2 F      alu = 2 W + 1.25 W_acc + 2 A      time = max(fma, alu)
// with W wide multiplies and the remaining instructions split into F (written as IMAD) and A (alu): what counts is
// the instruction total 3 W + F + A, split evenly.  ptxas balances the two pipes by COUNT (an IMAD.WIDE counts
// like an IADD3), turning movable additions into IMAD.X; the first b63 forms (6 W + 13.4 others: 32.8 cycles,
// 1133 G/s) left that to ptxas.  The forms below have fewer instructions (6 W + 10) and fix the split themselves:
//   * 64-bit accumulations run inside the multiplier, with its carry output where the sum can pass 2^64
//     (mul.wide + add.cc.u64 fuse into IMAD.WIDE.U32 Rd, P, ...; mul.wide + addc.u64 into IMAD.WIDE.U32.X);
//   * additions that must stay on the alu pipe are written so that no IMAD can express them: a consumed carry
//     out (the always-zero carry of a top limb feeds the carry INPUT of the next IMAD.WIDE), three register
//     operands, or two carry predicates (64-bit three-term sums become IADD3 Rd, P0, P1 / IADD3.X .., P0, P1);
//   * additions meant for the fma pipe are written as madc.lo (IMAD.X);
//   * the fold at bit 63 does not mask before it adds: top*c is added to the unmasked low 64 bits (mod 2^64) and
//     bit 63 is flipped afterwards if it was set (x - 2^63 = x xor 2^63 mod 2^64): one LOP3.
//   mul: 6 IMAD.WIDE, 1 IMAD, 1 IMAD.X, 8 alu: 30.05 cycles, 1239 G modmul/s (+9 %)
//   sqr: 5 IMAD.WIDE, 1 IMAD, 1 IMAD.X, 1 IMAD.SHL, 9 alu: 29.08 cycles, 1280 G/s (+8.6 %; all-alu: 33.6)
//   add / sub / neg: unchanged (1 IMAD + 7 / 9 / 6; ptxas moves two of the additions, which is the even split).
// Results: mul, sqr < 2^63 + 2^32; add < 2^63 + 2^16; sub < 2^63 + 2^17.
// c < 2^15 (all FireFly primes have c <= 13797), d = 2c.
// Word-level forms: a residue travels as two separate 32-bit registers (lo, hi).  The interpreter keeps
// its accumulators that way so that nothing but IMAD.WIDE's own result needs an aligned register pair --
// with 64-bit accumulators ptxas copied both halves of every accumulator at each tape instruction
// (16 IMAD.MOV per dispatch at 8 probes per thread, on the multiply pipe).
//
// x*y = t0 + 2^32 m + 2^64 t3  with t0 = x0 y0, m = x0 y1 + x1 y0 < 2^64, t3 = x1 y1 <= 2^62
//     = t0 + 2^32 ml + d h     (mod p),  h = t3 + mh < 2^62 + 2^32
//     = A + 2^32 (B + ml) + 2^64 cy,  A = (h0 d + t0) mod 2^64 with carry cy,  B = h1 d < 2^46
// V = that value (< 2^80), T = V >> 32 = A1 + B + ml + 2^32 cy, top = V >> 63 = T >> 31 (< 2^17):
// result = (V mod 2^63) + top c = ((V mod 2^64) + top c) xor (bit 63 of V), all mod 2^64.
#ifdef FF63_CY_ALU
#define FF63_CY "addc.u32 cy, 0, 0;\n\t"          /* SEL */
#else
#define FF63_CY "madc.lo.u32 cy, 0, 0, 0;\n\t"    /* IMAD.X */
#endif
#define FF63_MUL_REGS \
  ".reg .u32 ml,mh,tl,th,h0,h1,A0,A1,r1,top,cy,t,x;\n\t.reg .u64 t0,m,t3,A,B,U,U2,Z,T,X,hi;\n\t"
// expects t0, ml, and the carry chain h0 = ... (add.cc) just before it; C, DD name the operands c, d
#define FF63_MUL_TAIL(C, DD)                                                                          \
  "addc.cc.u32 h1, th, 0;\n\t"              /* carry out always 0 ... */                              \
  "mul.wide.u32 U2, h1, " DD ";\n\t"                                                                  \
  "addc.u64 B, U2, 0;\n\t"                  /* ... and consumed here: IMAD.WIDE.U32.X */              \
  "mul.wide.u32 U, h0, " DD ";\n\t"                                                                   \
  "add.cc.u64 A, U, t0;\n\t"                /* IMAD.WIDE.U32 with carry out */                        \
  FF63_CY                                                                                             \
  "shr.u64 hi, A, 32;\n\t"                                                                            \
  "mov.b64 Z, {ml,cy};\n\t"                                                                           \
  "add.u64 T, B, hi;\n\t"                                                                             \
  "add.u64 T, T, Z;\n\t"                    /* IADD3 with two carries out, IADD3.X with two carries in */ \
  "shr.u64 X, T, 31;\n\t"                                                                             \
  "cvt.u32.u64 top, X;\n\t"                                                                           \
  "cvt.u32.u64 r1, T;\n\t"                                                                            \
  "mov.b64 {A0,A1}, A;\n\t"                                                                           \
  "mul.lo.u32 t, top, " C ";\n\t"           /* < 2^32 */                                              \
  "add.cc.u32 %0, A0, t;\n\t"                                                                         \
  "madc.lo.u32 x, r1, 1, 0;\n\t"            /* IMAD.X: this one runs on the fma pipe */               \
  "lop3.b32 %1, x, r1, 0x80000000, 0x78;\n\t" /* x ^ (r1 & 0x80000000) */
FF63_D void mul_b63_w(u32& o0, u32& o1, u32 x0, u32 x1, u32 y0, u32 y1, u32 c, u32 d) {
  asm("{\n\t" FF63_MUL_REGS
      "mul.wide.u32 m, %2, %5;\n\t"
      "mad.wide.u32 m, %3, %4, m;\n\t"          // < 2^64: both high words are <= 2^31
      "mul.wide.u32 t0, %2, %4;\n\t"
      "mul.wide.u32 t3, %3, %5;\n\t"
      "mov.b64 {ml,mh}, m;\n\t"
      "mov.b64 {tl,th}, t3;\n\t"
      "add.cc.u32 h0, tl, mh;\n\t"
      FF63_MUL_TAIL("%6", "%7")
      "}" : "=r"(o0), "=r"(o1) : "r"(x0), "r"(x1), "r"(y0), "r"(y1), "r"(c), "r"(d));
}
FF63_D void sqr_b63_w(u32& o0, u32& o1, u32 x0, u32 x1, u32 c, u32 d) {
  asm("{\n\t" FF63_MUL_REGS
      "mul.wide.u32 m, %2, %3;\n\t"             // < 2^63
      "mul.wide.u32 t0, %2, %2;\n\t"
      "mul.wide.u32 t3, %3, %3;\n\t"
      "mov.b64 {ml,mh}, m;\n\t"
      "mov.b64 {tl,th}, t3;\n\t"
      "shf.l.wrap.b32 mh, ml, mh, 1;\n\t"       // the cross product counts twice: high word on the alu pipe,
      "mul.lo.u32 ml, ml, 2;\n\t"               // low word on the fma pipe (IMAD.SHL)
      "add.cc.u32 h0, tl, mh;\n\t"
      FF63_MUL_TAIL("%4", "%5")
      "}" : "=r"(o0), "=r"(o1) : "r"(x0), "r"(x1), "r"(c), "r"(d));
}
#ifdef FF63_ADD_ALU
// all-alu forms of add / sub (nothing ptxas can turn into IMAD.X), kept for the A/B: S = a + b, top = S >> 63,
// result = ((S mod 2^64) + top c) xor (bit 63 of S)
FF63_D void add_b63_w(u32& o0, u32& o1, u32 a0, u32 a1, u32 b0, u32 b1, u32 c) {
  asm("{\n\t"
      ".reg .u32 s0,s1,k,top,t,x;\n\t.reg .u64 a,b,tt,S;\n\t"
      "add.cc.u32 s0, %2, %4;\n\t"
      "addc.cc.u32 s1, %3, %5;\n\t"
      "addc.u32 k, 0, 0;\n\t"
      "shf.l.wrap.b32 top, s1, k, 1;\n\t"
      "mul.lo.u32 t, top, %6;\n\t"
      "mov.b64 a, {%2,%3};\n\t"
      "mov.b64 b, {%4,%5};\n\t"
      "cvt.u64.u32 tt, t;\n\t"
      "add.u64 S, a, b;\n\t"
      "add.u64 S, S, tt;\n\t"
      "mov.b64 {%0,x}, S;\n\t"
      "lop3.b32 %1, x, s1, 0x80000000, 0x78;\n\t"
      "}" : "=r"(o0), "=r"(o1) : "r"(a0), "r"(a1), "r"(b0), "r"(b1), "r"(c));
}
FF63_D void sub_b63_w(u32& o0, u32& o1, u32 a0, u32 a1, u32 b0, u32 b1, u32 c, u32 d) {
  asm("{\n\t"
      ".reg .u32 nd,m0,m1,s0,s1,k,top,t,x;\n\t.reg .u64 a,M,tt,S;\n\t"
      "neg.s32 nd, %7;\n\t"
      "sub.cc.u32 m0, nd, %4;\n\t"
      "subc.cc.u32 m1, 0xffffffff, %5;\n\t"
      "addc.cc.u32 s0, %2, m0;\n\t"
      "addc.cc.u32 s1, %3, m1;\n\t"
      "addc.u32 k, 0, 0;\n\t"
      "shf.l.wrap.b32 top, s1, k, 1;\n\t"
      "mul.lo.u32 t, top, %6;\n\t"
      "mov.b64 a, {%2,%3};\n\t"
      "mov.b64 M, {m0,m1};\n\t"
      "cvt.u64.u32 tt, t;\n\t"
      "add.u64 S, a, M;\n\t"
      "add.u64 S, S, tt;\n\t"
      "mov.b64 {%0,x}, S;\n\t"
      "lop3.b32 %1, x, s1, 0x80000000, 0x78;\n\t"
      "}" : "=r"(o0), "=r"(o1) : "r"(a0), "r"(a1), "r"(b0), "r"(b1), "r"(c), "r"(d));
}
#else
// a + b < 2^64 + 2^33: the carry and bit 63 together are worth (2k + b63) * c
FF63_D void add_b63_w(u32& o0, u32& o1, u32 a0, u32 a1, u32 b0, u32 b1, u32 c) {
  asm("{\n\t"
      ".reg .u32 s0,s1,k,top,t;\n\t"
      "add.cc.u32 s0, %2, %4;\n\t"
      "addc.cc.u32 s1, %3, %5;\n\t"
      "addc.u32 k, 0, 0;\n\t"
      "shf.l.wrap.b32 top, s1, k, 1;\n\t"       // (k << 1) | (s1 >> 31)  in {0, 1, 2}
      "and.b32 s1, s1, 0x7fffffff;\n\t"
      "mul.lo.u32 t, top, %6;\n\t"
      "add.cc.u32 %0, s0, t;\n\t"
      "addc.u32 %1, s1, 0;\n\t"
      "}" : "=r"(o0), "=r"(o1) : "r"(a0), "r"(a1), "r"(b0), "r"(b1), "r"(c));
}
// a - b = a + (2p - b); 2p = 2^64 - d.  2p - b < 2^64, the sum < 2^65: top in {0..3}
FF63_D void sub_b63_w(u32& o0, u32& o1, u32 a0, u32 a1, u32 b0, u32 b1, u32 c, u32 d) {
  asm("{\n\t"
      ".reg .u32 n0,m0,m1,s0,s1,k,top,t;\n\t"
      "neg.s32 n0, %7;\n\t"                     // low word of 2p
      "sub.cc.u32 m0, n0, %4;\n\t"
      "subc.u32 m1, 0xffffffff, %5;\n\t"
      "add.cc.u32 s0, %2, m0;\n\t"
      "addc.cc.u32 s1, %3, m1;\n\t"
      "addc.u32 k, 0, 0;\n\t"
      "shf.l.wrap.b32 top, s1, k, 1;\n\t"
      "and.b32 s1, s1, 0x7fffffff;\n\t"
      "mul.lo.u32 t, top, %6;\n\t"
      "add.cc.u32 %0, s0, t;\n\t"
      "addc.u32 %1, s1, 0;\n\t"
      "}" : "=r"(o0), "=r"(o1) : "r"(a0), "r"(a1), "r"(b0), "r"(b1), "r"(c), "r"(d));
}
#endif
// -a = 2p - a, folded at bit 63 (-0 gives 2p -> p, a lazy zero)
FF63_D void neg_b63_w(u32& o0, u32& o1, u32 a0, u32 a1, u32 c, u32 d) {
  asm("{\n\t"
      ".reg .u32 n0,s0,s1,top,t;\n\t"
      "neg.s32 n0, %5;\n\t"
      "sub.cc.u32 s0, n0, %2;\n\t"
      "subc.u32 s1, 0xffffffff, %3;\n\t"
      "shr.u32 top, s1, 31;\n\t"
      "and.b32 s1, s1, 0x7fffffff;\n\t"
      "mul.lo.u32 t, top, %4;\n\t"
      "add.cc.u32 %0, s0, t;\n\t"
      "addc.u32 %1, s1, 0;\n\t"
      "}" : "=r"(o0), "=r"(o1) : "r"(a0), "r"(a1), "r"(c), "r"(d));
}
// 64-bit wrappers
FF63_D u64 pack_w(u32 lo, u32 hi) { return ((u64)hi << 32) | lo; }
FF63_D u64 mul_b63(u64 x, u64 y, u32 c, u32 d) {
  u32 r0, r1;
  mul_b63_w(r0, r1, lo32(x), hi32(x), lo32(y), hi32(y), c, d);
  return pack_w(r0, r1);
}
FF63_D u64 sqr_b63(u64 x, u32 c, u32 d) {
  u32 r0, r1;
  sqr_b63_w(r0, r1, lo32(x), hi32(x), c, d);
  return pack_w(r0, r1);
}
FF63_D u64 add_b63(u64 a, u64 b, u32 c) {
  u32 r0, r1;
  add_b63_w(r0, r1, lo32(a), hi32(a), lo32(b), hi32(b), c);
  return pack_w(r0, r1);
}
FF63_D u64 sub_b63(u64 a, u64 b, u32 c, u32 d) {
  u32 r0, r1;
  sub_b63_w(r0, r1, lo32(a), hi32(a), lo32(b), hi32(b), c, d);
  return pack_w(r0, r1);
}
FF63_D u64 neg_b63(u64 a, u32 c, u32 d) {
  u32 r0, r1;
  neg_b63_w(r0, r1, lo32(a), hi32(a), c, d);
  return pack_w(r0, r1);
}

// branch-free lazy add / sub / neg (same values as add_lazy / sub_lazy / neg_lazy above)
FF63_D u64 add_lazy_dev(u64 a, u64 b, u32 d) {
  u64 r;
  asm("{\n\t"
      ".reg .u32 a0,a1,b0,b1,c,t;\n\t"
      "mov.b64 {a0,a1}, %1;\n\t"
      "mov.b64 {b0,b1}, %2;\n\t"
      "add.cc.u32 a0, a0, b0;\n\t"
      "addc.cc.u32 a1, a1, b1;\n\t"
      "addc.u32 c, 0, 0;\n\t"
      "mul.lo.u32 t, c, %3;\n\t"          // first wrap: + d
      "add.cc.u32 a0, a0, t;\n\t"
      "addc.cc.u32 a1, a1, 0;\n\t"
      "addc.u32 c, 0, 0;\n\t"
      "mad.lo.u32 a0, c, %3, a0;\n\t"     // second wrap: value < d now, + d cannot carry
      "mov.b64 %0, {a0,a1};\n\t"
      "}" : "=l"(r) : "l"(a), "l"(b), "r"(d));
  return r;
}
FF63_D u64 sub_lazy_dev(u64 a, u64 b, u32 d, u64 p) {
  u64 r;
  asm("{\n\t"
      ".reg .u32 a0,a1,b0,b1,p0,p1,m,t;\n\t"
      "mov.b64 {a0,a1}, %1;\n\t"
      "mov.b64 {b0,b1}, %2;\n\t"
      "mov.b64 {p0,p1}, %4;\n\t"
      "sub.cc.u32 a0, a0, b0;\n\t"
      "subc.cc.u32 a1, a1, b1;\n\t"
      "subc.u32 m, 0, 0;\n\t"             // m = 0xffffffff on borrow
      "and.b32 t, m, %3;\n\t"
      "sub.cc.u32 a0, a0, t;\n\t"         // borrow: - d
      "subc.cc.u32 a1, a1, 0;\n\t"
      "subc.u32 m, 0, 0;\n\t"
      "and.b32 p0, p0, m;\n\t"
      "and.b32 p1, p1, m;\n\t"
      "add.cc.u32 a0, a0, p0;\n\t"        // went below zero: + p
      "addc.u32 a1, a1, p1;\n\t"
      "mov.b64 %0, {a0,a1};\n\t"
      "}" : "=l"(r) : "l"(a), "l"(b), "r"(d), "l"(p));
  return r;
}
FF63_D u64 neg_lazy_dev(u64 a, u32 d, u64 p) { return sub_lazy_dev(0, a, d, p); }
#endif  // __CUDACC__

}  // namespace ff63
