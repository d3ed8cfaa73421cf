// gmpxx.h -- TEST INFRASTRUCTURE ONLY.  Minimal stand-in for GMP's C++ header.
//
// The authoring container and the GPU boxes ship the GMP *runtime*
// (libgmp.so.10) but neither gmp.h nor gmpxx.h, so the reference
// (jklappert/FireFly, which includes <gmpxx.h> from FFInt.hpp:27) cannot be
// compiled as shipped.  This header declares just the libgmp entry points the
// reference uses and a small value-semantic mpz_class on top of them, so that
// the UNMODIFIED reference sources can be built into oracle/_ref/ as the parity
// checker.  It is not part of the product and is never linked into it.
#pragma once
#include <algorithm>
#include <cstddef>
#include <cstdlib>
#include <iostream>
#include <stdexcept>
#include <string>

extern "C" {
typedef unsigned long mp_limb_t;
typedef struct {
  int _mp_alloc;
  int _mp_size;
  mp_limb_t* _mp_d;
} __mpz_struct;
typedef __mpz_struct mpz_t[1];
typedef __mpz_struct* mpz_ptr;
typedef const __mpz_struct* mpz_srcptr;

void __gmpz_init(mpz_ptr);
void __gmpz_clear(mpz_ptr);
void __gmpz_init_set(mpz_ptr, mpz_srcptr);
int __gmpz_init_set_str(mpz_ptr, const char*, int);
void __gmpz_init_set_si(mpz_ptr, long);
void __gmpz_init_set_ui(mpz_ptr, unsigned long);
void __gmpz_set(mpz_ptr, mpz_srcptr);
int __gmpz_set_str(mpz_ptr, const char*, int);
void __gmpz_add(mpz_ptr, mpz_srcptr, mpz_srcptr);
void __gmpz_sub(mpz_ptr, mpz_srcptr, mpz_srcptr);
void __gmpz_mul(mpz_ptr, mpz_srcptr, mpz_srcptr);
void __gmpz_tdiv_q(mpz_ptr, mpz_srcptr, mpz_srcptr);
void __gmpz_tdiv_r(mpz_ptr, mpz_srcptr, mpz_srcptr);
void __gmpz_neg(mpz_ptr, mpz_srcptr);
void __gmpz_abs(mpz_ptr, mpz_srcptr);
void __gmpz_gcd(mpz_ptr, mpz_srcptr, mpz_srcptr);
int __gmpz_invert(mpz_ptr, mpz_srcptr, mpz_srcptr);
size_t __gmpz_sizeinbase(mpz_srcptr, int);
int __gmpz_cmp(mpz_srcptr, mpz_srcptr);
int __gmpz_cmp_si(mpz_srcptr, long);
char* __gmpz_get_str(char*, int, mpz_srcptr);
}

#define mpz_init __gmpz_init
#define mpz_clear __gmpz_clear
#define mpz_invert __gmpz_invert
#define mpz_sizeinbase __gmpz_sizeinbase

class mpz_class {
 public:
  mpz_class() { __gmpz_init(v_); }
  mpz_class(const mpz_class& o) { __gmpz_init_set(v_, o.v_); }
  mpz_class(const mpz_t o) { __gmpz_init_set(v_, o); }
  mpz_class(int x) { __gmpz_init_set_si(v_, x); }
  mpz_class(long x) { __gmpz_init_set_si(v_, x); }
  mpz_class(long long x) { __gmpz_init_set_si(v_, (long)x); }
  mpz_class(unsigned x) { __gmpz_init_set_ui(v_, x); }
  mpz_class(unsigned long x) { __gmpz_init_set_ui(v_, x); }
  mpz_class(unsigned long long x) { __gmpz_init_set_ui(v_, (unsigned long)x); }
  // real gmpxx defaults to base 0 (0x / 0b / leading-0 octal prefixes are honoured)
  explicit mpz_class(const std::string& s, int base = 0) { init_str(s.c_str(), base); }
  explicit mpz_class(const char* s, int base = 0) { init_str(s, base); }
  ~mpz_class() { __gmpz_clear(v_); }
  mpz_class& operator=(const mpz_class& o) {
    if (this != &o) __gmpz_set(v_, o.v_);
    return *this;
  }
  mpz_class& operator=(const std::string& s) {
    __gmpz_set_str(v_, s.c_str(), 0);
    return *this;
  }
  mpz_ptr get_mpz_t() { return v_; }
  mpz_srcptr get_mpz_t() const { return v_; }
  std::string get_str(int base = 10) const {
    char* c = __gmpz_get_str(nullptr, base, v_);
    std::string s(c);
    std::free(c);
    return s;
  }
  mpz_class& operator+=(const mpz_class& o) { __gmpz_add(v_, v_, o.v_); return *this; }
  mpz_class& operator-=(const mpz_class& o) { __gmpz_sub(v_, v_, o.v_); return *this; }
  mpz_class& operator*=(const mpz_class& o) { __gmpz_mul(v_, v_, o.v_); return *this; }
  mpz_class& operator/=(const mpz_class& o) { __gmpz_tdiv_q(v_, v_, o.v_); return *this; }
  mpz_class& operator%=(const mpz_class& o) { __gmpz_tdiv_r(v_, v_, o.v_); return *this; }
  mpz_class operator-() const { mpz_class r; __gmpz_neg(r.v_, v_); return r; }
  mpz_class operator+() const { return *this; }
  friend mpz_class operator+(const mpz_class& a, const mpz_class& b) { mpz_class r; __gmpz_add(r.v_, a.v_, b.v_); return r; }
  friend mpz_class operator-(const mpz_class& a, const mpz_class& b) { mpz_class r; __gmpz_sub(r.v_, a.v_, b.v_); return r; }
  friend mpz_class operator*(const mpz_class& a, const mpz_class& b) { mpz_class r; __gmpz_mul(r.v_, a.v_, b.v_); return r; }
  friend mpz_class operator/(const mpz_class& a, const mpz_class& b) { mpz_class r; __gmpz_tdiv_q(r.v_, a.v_, b.v_); return r; }
  friend mpz_class operator%(const mpz_class& a, const mpz_class& b) { mpz_class r; __gmpz_tdiv_r(r.v_, a.v_, b.v_); return r; }
  friend bool operator==(const mpz_class& a, const mpz_class& b) { return __gmpz_cmp(a.v_, b.v_) == 0; }
  friend bool operator!=(const mpz_class& a, const mpz_class& b) { return __gmpz_cmp(a.v_, b.v_) != 0; }
  friend bool operator<(const mpz_class& a, const mpz_class& b) { return __gmpz_cmp(a.v_, b.v_) < 0; }
  friend bool operator<=(const mpz_class& a, const mpz_class& b) { return __gmpz_cmp(a.v_, b.v_) <= 0; }
  friend bool operator>(const mpz_class& a, const mpz_class& b) { return __gmpz_cmp(a.v_, b.v_) > 0; }
  friend bool operator>=(const mpz_class& a, const mpz_class& b) { return __gmpz_cmp(a.v_, b.v_) >= 0; }
  friend std::ostream& operator<<(std::ostream& o, const mpz_class& a) { return o << a.get_str(); }
  friend mpz_class abs(const mpz_class& a) { mpz_class r; __gmpz_abs(r.v_, a.v_); return r; }
  friend mpz_class gcd(const mpz_class& a, const mpz_class& b) { mpz_class r; __gmpz_gcd(r.v_, a.v_, b.v_); return r; }

 private:
  void init_str(const char* s, int base) {
    if (__gmpz_init_set_str(v_, s, base) != 0) {
      __gmpz_clear(v_);
      throw std::invalid_argument("mpz_class: bad integer string");
    }
  }
  mpz_t v_;
};
