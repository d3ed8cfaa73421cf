// ffhost.hpp -- scalar Z_p arithmetic on the HOST with the reference's FFInt
// semantics.  Used by the tape compiler only: to reduce literals, fold constant
// sub-expressions once per prime and to fingerprint functions for duplicate
// detection.  Probe values are never computed here (there is no CPU path for
// the hot loop); this is the analogue of what ShuntingYardParser::precompute
// does with FFInt on the parse side (reference: source/ShuntingYardParser.cpp:593-718).
//
// Semantics reproduced (reference DEFAULT arithmetic, source/FFInt.cpp:130-192,
// 226-228, 356-405 and source/include/firefly/FFInt.hpp:129-174):
//   a/0 = 0, -0 = 0, x^0 = 1 (also 0^0), exponent residue >= p/2 means a negative
//   power, to_neg_int truncates to 32 bits.
#pragma once
#include <cstdint>
#include <string>

namespace ffhost {

typedef unsigned __int128 u128;

inline uint64_t add(uint64_t a, uint64_t b, uint64_t p) {
  uint64_t s = a + b;                       // p < 2^63: no wrap
  return s >= p ? s - p : s;
}
inline uint64_t sub(uint64_t a, uint64_t b, uint64_t p) { return a >= b ? a - b : a + p - b; }
inline uint64_t neg(uint64_t a, uint64_t p) { return a ? p - a : 0; }
inline uint64_t mul(uint64_t a, uint64_t b, uint64_t p) { return (uint64_t)((u128)a * b % p); }

// modular inverse with inv(0) = 0 (reference mod_inv leaves t = 0 when a = 0, FFInt.cpp:385-405)
inline uint64_t inv(uint64_t a, uint64_t p) {
  __int128 t = 0, nt = 1;
  uint64_t r = p, nr = a % p;
  while (nr) {
    uint64_t q = r / nr;
    __int128 tt = t - (__int128)q * nt; t = nt; nt = tt;
    uint64_t rr = r - q * nr; r = nr; nr = rr;
  }
  if (t < 0) t += p;
  return (uint64_t)t;
}
// FFInt::operator/ : zero divisor gives zero (FFInt.cpp:146-153)
inline uint64_t div(uint64_t a, uint64_t b, uint64_t p) { return b == 0 ? 0 : mul(a, inv(b, p), p); }

// mod_pow: x^0 = 1 for every x (FFInt.cpp:371-383)
inline uint64_t pow_u(uint64_t b, uint64_t e, uint64_t p) {
  uint64_t r = 1;
  while (e) {
    if (e & 1) r = mul(r, b, p);
    b = mul(b, b, p);
    e >>= 1;
  }
  return r;
}
// FFInt::pow(const FFInt&) (FFInt.cpp:155-179): exponent is a residue
inline uint64_t pow_ff(uint64_t b, uint64_t e, uint64_t p) {
  if (e == 2) return mul(b, b, p);
  if (e < (p >> 1)) return pow_u(b, e, p);
  return pow_u(inv(b, p), p - e - 1, p);
}
// FFInt::to_neg_int (FFInt.cpp:226-228)
inline int32_t to_neg_int(uint64_t n) { return (int32_t)(0u - (uint32_t)n); }
// FFInt::pow<int> (FFInt.hpp:149-174)
inline uint64_t pow_int(uint64_t b, int32_t e, uint64_t p) {
  if (e == 2) return mul(b, b, p);
  if (e >= 0) return pow_u(b, (uint64_t)e, p);
  return pow_u(inv(b, p), (uint64_t)(-(int64_t)e), p);
}

// decimal digits (no sign) -> residue; any length
inline uint64_t reduce_decimal(const char* s, size_t n, uint64_t p) {
  uint64_t r = 0;
  for (size_t i = 0; i < n; ++i) r = (uint64_t)(((u128)r * 10 + (unsigned)(s[i] - '0')) % p);
  return r;
}

}  // namespace ffhost
