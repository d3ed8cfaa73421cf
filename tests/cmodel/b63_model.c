// b63_model.c -- word-level C model of the device arithmetic (firefly_b200/csrc/ff63.cuh: mul_b63_w / sqr_b63_w,
// add_b63_w / sub_b63_w / neg_b63_w, canon_pm): the same steps with explicit widths, every "cannot overflow" claim of the PTX as an
// assert, the result compared with (x*y) mod p from unsigned __int128 on edge and random operands below the
// representation bound 2^63 + 2^32, including operands that force the carry output of the multiplier.
// Test infrastructure (tests/test_arith_model.py compiles and runs it); not part of the product.
// usage: b63_model [random cases per prime]
#include <stdio.h>
#include <stdint.h>
#include <stdlib.h>
#include <assert.h>
typedef unsigned __int128 u128; typedef uint64_t u64; typedef uint32_t u32;
static u64 P, Bnd = (1ULL<<63) + (1ULL<<32);
static u32 c, d;
static long n_carry = 0;
static u64 W(u32 a, u32 b, u64 C, int cin, int* cout) { u128 s = (u128)a*b + C + cin; if (cout) *cout = (int)(s >> 64); else assert((s >> 64) == 0); return (u64)s; }
static u64 mulN(u64 x, u64 y) {
  u32 x0 = (u32)x, x1 = x >> 32, y0 = (u32)y, y1 = y >> 32;
  u64 m = W(x0, y1, 0, 0, 0); m = W(x1, y0, m, 0, 0);
  u64 t0 = W(x0, y0, 0, 0, 0), t3 = W(x1, y1, 0, 0, 0);
  u32 ml = (u32)m, mh = m >> 32;
  u64 h = t3 + mh; assert(h >= t3);
  u32 h0 = (u32)h, h1 = h >> 32;
  int cA; u64 A = W(h0, d, t0, 0, &cA);
  u64 Bv = W(h1, d, 0, 0, 0);
  u32 A0 = (u32)A, A1 = A >> 32;
  u128 R = (u128)Bv + A1 + ml + ((u64)cA << 32); assert((R >> 64) == 0);
  u32 r1 = (u32)R, r2 = (u32)(R >> 32);
  u32 top = (r2 << 1) | (r1 >> 31); assert(r2 < (1u << 31));
  // shipped form: add top*c to the UNMASKED low 64 bits mod 2^64, then flip bit 63 if it was set
  u32 t = top * c; assert((u64)top * c < (1ULL << 32));
  u32 o0 = A0 + t; u32 cyo = o0 < A0;
  u32 xh = r1 + cyo;                                  // madc.lo x, r1, 1, 0
  u32 o1 = xh ^ (r1 & 0x80000000u);
  u64 res = ((u64)o1 << 32) | o0;
  // reference form: mask, then add
  u64 ref = W(top, c, ((u64)(r1 & 0x7fffffff) << 32) | A0, 0, 0);
  assert(res == ref);
  assert(res < Bnd);
  n_carry += cA;
  return res;
}
static u64 sqrN(u64 x) {
  u32 x0 = (u32)x, x1 = x >> 32;
  u64 m = W(x0, x1, 0, 0, 0); assert(m < (1ULL << 63));
  u64 t0 = W(x0, x0, 0, 0, 0), t3 = W(x1, x1, 0, 0, 0);
  m <<= 1;
  u32 ml = (u32)m, mh = m >> 32;
  u64 h = t3 + mh; assert(h >= t3);
  u32 h0 = (u32)h, h1 = h >> 32;
  int cA; u64 A = W(h0, d, t0, 0, &cA);
  u64 Bv = W(h1, d, 0, 0, 0);
  u32 A0 = (u32)A, A1 = A >> 32;
  u128 R = (u128)Bv + A1 + ml + ((u64)cA << 32); assert((R >> 64) == 0);
  u32 r1 = (u32)R, r2 = (u32)(R >> 32);
  u32 top = (r2 << 1) | (r1 >> 31); assert(r2 < (1u << 31));
  // shipped form: add top*c to the UNMASKED low 64 bits mod 2^64, then flip bit 63 if it was set
  u32 t = top * c; assert((u64)top * c < (1ULL << 32));
  u32 o0 = A0 + t; u32 cyo = o0 < A0;
  u32 xh = r1 + cyo;                                  // madc.lo x, r1, 1, 0
  u32 o1 = xh ^ (r1 & 0x80000000u);
  u64 res = ((u64)o1 << 32) | o0;
  // reference form: mask, then add
  u64 ref = W(top, c, ((u64)(r1 & 0x7fffffff) << 32) | A0, 0, 0);
  assert(res == ref);
  assert(res < Bnd);
  n_carry += cA;
  return res;
}
// add_b63_w: s = a + b with carry k; the carry and bit 63 together are worth (2k + b63) * c
static u64 addN(u64 a, u64 b) {
  u32 a0 = (u32)a, a1 = a >> 32, b0 = (u32)b, b1 = b >> 32;
  u64 lo = (u64)a0 + b0; u32 s0 = (u32)lo;
  u64 hi = (u64)a1 + b1 + (lo >> 32); u32 s1 = (u32)hi; u32 k = (u32)(hi >> 32);
  u32 top = (k << 1) | (s1 >> 31); assert(top <= 2);
  s1 &= 0x7fffffffu;
  u32 t = top * c; assert((u64)top * c < (1ULL << 32));
  u64 r0 = (u64)s0 + t; u64 r1 = (u64)s1 + (r0 >> 32); assert((r1 >> 32) == 0);
  u64 res = (r1 << 32) | (u32)r0;
  assert(res < (1ULL << 63) + (1ULL << 16));               // the bound ff63.cuh states
  assert(res < Bnd);
  return res;
}
// sub_b63_w: a - b = a + (2p - b), 2p = 2^64 - d; the sum is below 2^65: top in {0..3}
static u64 subN(u64 a, u64 b) {
  u32 a0 = (u32)a, a1 = a >> 32, b0 = (u32)b, b1 = b >> 32;
  u32 n0 = 0u - d;                                         // low word of 2p
  u64 m = (((u64)0xffffffffu << 32) | n0) - b;             // sub.cc / subc: no borrow out, b <= 2p
  assert(b <= (((u64)0xffffffffu << 32) | n0));
  u32 m0 = (u32)m, m1 = m >> 32; (void)b0; (void)b1;
  u64 lo = (u64)a0 + m0; u32 s0 = (u32)lo;
  u64 hi = (u64)a1 + m1 + (lo >> 32); u32 s1 = (u32)hi; u32 k = (u32)(hi >> 32);
  u32 top = (k << 1) | (s1 >> 31); assert(top <= 3);
  s1 &= 0x7fffffffu;
  u32 t = top * c;
  u64 r0 = (u64)s0 + t; u64 r1 = (u64)s1 + (r0 >> 32); assert((r1 >> 32) == 0);
  u64 res = (r1 << 32) | (u32)r0;
  assert(res < (1ULL << 63) + (1ULL << 17));
  assert(res < Bnd);
  return res;
}
// neg_b63_w: -a = 2p - a, folded at bit 63
static u64 negN(u64 a) {
  u32 n0 = 0u - d;
  u64 s = (((u64)0xffffffffu << 32) | n0) - a;
  u32 s0 = (u32)s, s1 = s >> 32;
  u32 top = s1 >> 31;
  s1 &= 0x7fffffffu;
  u32 t = top * c;
  u64 r0 = (u64)s0 + t; u64 r1 = (u64)s1 + (r0 >> 32); assert((r1 >> 32) == 0);
  u64 res = (r1 << 32) | (u32)r0;
  assert(res < Bnd);
  return res;
}
// canon_pm: v = b * 2^63 + l  ->  l + b * c < 2^63 + c < 2p, one conditional subtraction
static u64 canonN(u64 v) {
  u64 l = (v & 0x7fffffffffffffffULL) + ((v >> 63) ? (u64)c : 0ULL);
  assert(l < 2 * P);
  u64 r = l >= P ? l - P : l;
  assert(r < P && r == v % P);
  return r;
}
// inv_fermat (kernels.cu): x^(p-2), p - 2 = (2^47 - 1) * 2^16 + (2^16 - c - 2); 2^47 - 1 by the addition chain
// 1,2,3,5,10,20,23,46,47 on the run length, then 16 squarings interleaved with the bits of the tail
static u64 sqr_n(u64 x, int n) { while (n--) x = sqrN(x); return x; }
static u64 invN(u64 x) {
  const u64 x1 = x;
  const u64 x2 = mulN(sqrN(x1), x1);
  const u64 x3 = mulN(sqrN(x2), x1);
  const u64 x5 = mulN(sqr_n(x3, 2), x2);
  const u64 x10 = mulN(sqr_n(x5, 5), x5);
  const u64 x20 = mulN(sqr_n(x10, 10), x10);
  const u64 x23 = mulN(sqr_n(x20, 3), x3);
  const u64 x46 = mulN(sqr_n(x23, 23), x23);
  u64 r = mulN(sqrN(x46), x1);
  const u32 tail = 65536u - c - 2u;
  for (int bit = 15; bit >= 0; --bit) {
    r = sqrN(r);
    if ((tail >> bit) & 1) r = mulN(r, x1);
  }
  return r;
}
static u64 rnd(void) { static u64 s = 88172645463325252ULL; s ^= s << 13; s ^= s >> 7; s ^= s << 17; return s; }
int main(int argc, char** argv) {
  const long n_random = argc > 1 ? atol(argv[1]) : 2000000;
  u32 cs[] = {25, 165, 13797, 1, 32767};     // the first three are primes of FireFly's table, the others stress the widths
  long n = 0, ncA = 0;
  for (int ci = 0; ci < 5; ci++) {
    c = cs[ci]; d = 2*c; P = (1ULL<<63) - c;
    u64 edge[] = {0,1,2,P-1,P,P+1,(1ULL<<63)-1,1ULL<<63,(1ULL<<63)+1,Bnd-1,Bnd-2,0xffffffffULL,0x100000000ULL,0x7fffffffffffffffULL,
                  0x80000000ffffffffULL,0x7fffffff00000000ULL,0x7fffffffffff0000ULL,0xfffffffeULL,0x80000000fffffffeULL,0x7ffffffffffffffeULL, P-2, (P+1)/2, 0x00000000ffff0001ULL};
    int ne = sizeof(edge)/sizeof(edge[0]);
    for (int i = 0; i < ne; i++) for (int j = 0; j < ne; j++) {
      u64 r = mulN(edge[i], edge[j]); assert(r % P == (u64)(((u128)edge[i] * edge[j]) % P));
      if (i == j) { u64 s = sqrN(edge[i]); assert(s % P == r % P); }
      const u64 x = edge[i], y = edge[j];
      assert(addN(x, y) % P == (u64)(((u128)x + y) % P));
      assert(subN(x, y) % P == (u64)(((u128)x + 2 * (u128)P - y) % P));
      if (i == j) {
        assert(negN(x) % P == (u64)((2 * (u128)P - x) % P)); canonN(x); canonN(r);
        const u64 iv = ci < 3 ? invN(x) : 0;                 // 0 -> 0, else x * x^(p-2) = 1 (the prime moduli only)
        assert(ci >= 3 || (u64)(((u128)(iv % P) * (x % P)) % P) == (x % P ? 1 : 0));
      }
      n++;
    }
    for (long k = 0; k < n_random; k++) {
      u64 x = rnd(), y = rnd();
      int mode = k & 7;
      if (mode == 1) { x |= 0xffff0000u; y |= 0xffff0000u; }          // low words near 2^32: exercises cA
      if (mode == 2) { x |= 0xfffffff0u; y |= 0xfffffffeu; x |= 0x7fff000000000000ULL; y |= 0x7fffffff00000000ULL; }
      if (mode == 3) { x = (x & 0xffffffffULL) | (1ULL << 63); }
      x %= Bnd; y %= Bnd;
      u64 r = mulN(x, y); assert(r % P == (u64)(((u128)x * y) % P));
      u64 s = sqrN(x); assert(s % P == (u64)(((u128)x * x) % P));
      // additions and subtractions of results of every kind stay inside the representation
      u64 ad = addN(x, y), sb = subN(x, y), ng = negN(x);
      assert(ad % P == (u64)(((u128)x + y) % P));
      assert(sb % P == (u64)(((u128)x + 2 * (u128)P - y) % P));
      assert(ng % P == (u64)((2 * (u128)P - x) % P));
      assert(addN(ad, sb) % P == (u64)(((u128)ad + sb) % P));
      assert(subN(ng, r) % P == (u64)(((u128)ng + 2 * (u128)P - r) % P));
      assert(mulN(ad, sb) % P == (u64)(((u128)ad * sb) % P));
      canonN(ad); canonN(sb); canonN(ng); canonN(r); canonN(s);
      if (ci < 3 && (k & 1023) == 0) { const u64 iv = invN(x); assert((u64)(((u128)(iv % P) * (x % P)) % P) == (x % P ? 1 : 0)); }
      n++;
    }
  }
  printf("ok %ld cases, multiplier carry taken %ld times\n", n, n_carry);
  if (n_carry == 0) return 3;
  return 0;
}
