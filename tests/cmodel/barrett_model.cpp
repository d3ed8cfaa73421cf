// Host check of ff63::mul_barrett (firefly_b200/csrc/ff63.cuh) -- the multiplication of the generic-modulus kernels,
// the same source the device compiles -- against unsigned __int128 for moduli of every size below 2^63.
// TEST INFRASTRUCTURE (tests/test_arith_model.py).  Reference semantics: FFInt::operator*= / mod_mul return the
// canonical residue of the exact product for ANY modulus (reference: source/FFInt.cpp:141-144, 356-369).
#include <cstdint>
#include <cstdio>
#include <cstdlib>

#include "../../firefly_b200/csrc/ff63.cuh"

static uint64_t s = 0x9E3779B97F4A7C15ULL;
static uint64_t rnd() {
  uint64_t z = (s += 0x9E3779B97F4A7C15ULL);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
  return z ^ (z >> 31);
}

int main(int argc, char** argv) {
  const long n = argc > 1 ? atol(argv[1]) : 200000;
  // odd moduli: tiny, every bit length, next to the powers of two, FireFly's largest primes, random
  uint64_t mods[400];
  int nm = 0;
  const uint64_t fixed[] = {3, 5, 7, 101, 65537, 4294967291ULL, 4294967311ULL, 2305843009213693951ULL, 4611686018427387847ULL,
                            4611686018427387903ULL, 4611686018427387905ULL, 9223372036854775783ULL, 9223372036854775643ULL,
                            9223372036854775807ULL, 9223372036854742951ULL, 9223372036854710017ULL, 6917529027641081903ULL};
  for (uint64_t m : fixed) mods[nm++] = m;
  for (int bits = 2; bits <= 63; ++bits)
    for (int r = 0; r < 4; ++r) {
      uint64_t m = (rnd() >> (64 - bits)) | (1ULL << (bits - 1)) | 1ULL;
      mods[nm++] = m;
    }
  long checked = 0;
  for (int i = 0; i < nm; ++i) {
    const uint64_t p = mods[i];
    const ff63::Barrett f = ff63::make_barrett(p);
    const uint64_t edge[] = {0, 1, 2, p - 1, p - 2, p >> 1, (p >> 1) + 1};
    for (uint64_t a : edge)
      for (uint64_t b : edge) {
        if (a >= p || b >= p) continue;
        const uint64_t want = (uint64_t)((unsigned __int128)a * b % p);
        if (ff63::mul_barrett(a, b, f) != want) { printf("FAIL p=%llu a=%llu b=%llu\n", (unsigned long long)p, (unsigned long long)a, (unsigned long long)b); return 1; }
        ++checked;
      }
    for (long k = 0; k < n / nm + 1; ++k) {
      const uint64_t a = rnd() % p, b = (k & 7) ? rnd() % p : p - 1 - (rnd() % 3 % p);
      const uint64_t want = (uint64_t)((unsigned __int128)a * b % p);
      if (ff63::mul_barrett(a, b, f) != want) { printf("FAIL p=%llu a=%llu b=%llu\n", (unsigned long long)p, (unsigned long long)a, (unsigned long long)b); return 1; }
      ++checked;
    }
  }
  printf("ok %ld products over %d moduli\n", checked, nm);
  return 0;
}
