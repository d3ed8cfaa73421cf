/* ff_oracle.c -- TEST INFRASTRUCTURE ONLY (see ff_oracle.h).
 *
 * A plain-C restatement of the CPU algorithm the GPU path must reproduce bit for bit:
 * jklappert/FireFly 2.0.3, ShuntingYardParser (text -> RPN -> precomputed tokens ->
 * stack-machine evaluation) over FFInt's DEFAULT arithmetic.  Deliberately kept in the
 * reference's own shape -- a flat (opcode, value) token list per function walked by a stack
 * machine -- so that it checks the product's very different pipeline (expression trees,
 * accumulator tape, lazy residues, batched inversion) independently.
 *
 * Pinned against the unmodified reference (oracle/_ref) by tests/test_oracle.py and against
 * the JSON files in tests/golden.  Nothing under firefly_b200/ links, loads or calls this file.
 */
#include "ff_oracle.h"

#include <ctype.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

typedef unsigned __int128 u128;

static _Thread_local char g_err[512];
const char* ffo_last_error(void) { return g_err; }
static void set_err(const char* a, const char* b, const char* c) {
  snprintf(g_err, sizeof g_err, "%s%s%s", a, b ? b : "", c ? c : "");
}

/* ------------------------------------------------------------------------------------------ */
/* field arithmetic                                                                            */
/* ------------------------------------------------------------------------------------------ */

/* FFInt::operator+= (FFInt.cpp:131-134): n = (p - b > a) ? a + b : a + b - p */
uint64_t ffo_add(uint64_t a, uint64_t b, uint64_t p) { return (p - b > a) ? a + b : a + b - p; }
/* FFInt::operator-= (FFInt.cpp:136-139) */
uint64_t ffo_sub(uint64_t a, uint64_t b, uint64_t p) { return (b > a) ? a - b + p : a - b; }
/* FFInt::operator-() (FFInt.cpp:190-192) builds FFInt(p - n); the constructor (FFInt.hpp:129-140)
 * reduces p to 0, so -0 = 0 */
uint64_t ffo_neg(uint64_t a, uint64_t p) {
  uint64_t v = p - a;
  return v < p ? v : v % p;
}
/* mod_mul (FFInt.cpp:356-369) returns the exact product residue; the reference gets there with a
 * long-double quotient estimate, here with a 128-bit product -- same value */
uint64_t ffo_mul(uint64_t a, uint64_t b, uint64_t p) {
  if (a >= p) a %= p;
  if (b >= p) b %= p;
  return (uint64_t)((u128)a * b % p);
}
/* mod_inv (FFInt.cpp:385-405): extended Euclid; a = 0 leaves t = 0 */
uint64_t ffo_inv(uint64_t a, uint64_t p) {
  __int128 t = 0, newt = 1;
  uint64_t r = p, newr = a;
  while (newr) {
    uint64_t q = r / newr;
    __int128 tmpt = t;
    t = newt;
    newt = tmpt - (__int128)q * newt;
    uint64_t tmpr = r;
    r = newr;
    newr = tmpr - q * newr;
  }
  return (uint64_t)(t < 0 ? t + (__int128)p : t);
}
/* FFInt::operator/= (FFInt.cpp:146-153): zero divisor -> 0 */
uint64_t ffo_div(uint64_t a, uint64_t b, uint64_t p) {
  if (b == 0) return 0;
  return ffo_mul(a, ffo_inv(b, p), p);
}
/* mod_pow (FFInt.cpp:371-383) */
static uint64_t mod_pow(uint64_t base, uint64_t e, uint64_t p) {
  uint64_t res = 1;
  while (e) {
    if (e & 1) res = ffo_mul(res, base, p);
    base = ffo_mul(base, base, p);
    e >>= 1;
  }
  return res;
}
/* FFInt::pow(const FFInt&) (FFInt.cpp:155-179) */
uint64_t ffo_pow(uint64_t a, uint64_t e, uint64_t p) {
  if (e == 2) return ffo_mul(a, a, p);
  if (e < (p >> 1)) return mod_pow(a, e, p);
  return mod_pow(ffo_inv(a, p), p - e - 1, p);
}
/* FFInt::to_neg_int (FFInt.cpp:226-228): -(int)n */
int32_t ffo_to_neg_int(uint64_t a) { return (int32_t)(0u - (uint32_t)a); }
/* FFInt::pow<int> (FFInt.hpp:149-174) */
uint64_t ffo_pow_int(uint64_t a, int32_t e, uint64_t p) {
  if (e == 2) return ffo_mul(a, a, p);
  if (e >= 0) return mod_pow(a, (uint64_t)e, p);
  return mod_pow(ffo_inv(a, p), (uint64_t)(-(int64_t)e), p);
}

/* ------------------------------------------------------------------------------------------ */
/* prime table: the 300 largest primes below 2^63 (ReconstHelper.cpp:24-125), recomputed        */
/* ------------------------------------------------------------------------------------------ */
static int is_prime(uint64_t n) {
  static const uint64_t small[] = {2, 3, 5, 7, 11, 13, 17, 19, 23, 29, 31, 37};
  for (int i = 0; i < 12; i++)
    if (n % small[i] == 0) return n == small[i];
  uint64_t d = n - 1;
  int s = 0;
  while (!(d & 1)) { d >>= 1; s++; }
  for (int i = 0; i < 12; i++) {
    uint64_t x = mod_pow(small[i], d, n);
    if (x == 1 || x == n - 1) continue;
    int composite = 1;
    for (int r = 1; r < s; r++) {
      x = ffo_mul(x, x, n);
      if (x == n - 1) { composite = 0; break; }
    }
    if (composite) return 0;
  }
  return 1;
}
static uint64_t g_primes[300];
static int g_primes_ready;
static void fill_primes(void) {
  if (g_primes_ready) return;
  int k = 0;
  for (uint64_t n = ((uint64_t)1 << 63) - 1; k < 300; n -= 2)
    if (is_prime(n)) g_primes[k++] = n;
  g_primes_ready = 1;
}
uint32_t ffo_n_primes(void) { return 300; }
uint64_t ffo_prime(uint32_t i) {
  fill_primes();
  return i < 300 ? g_primes[i] : 0;
}

/* ------------------------------------------------------------------------------------------ */
/* bunch policy (utils.cpp:305-344)                                                            */
/* ------------------------------------------------------------------------------------------ */
uint32_t ffo_compute_bunch_size(uint32_t queue_length, uint32_t thr_n, uint32_t max_bunch_size) {
  if (max_bunch_size == 1) return 1;
  uint32_t tmp = queue_length / thr_n;
  if (tmp == 0) return 1;
  tmp |= tmp >> 1; tmp |= tmp >> 2; tmp |= tmp >> 4; tmp |= tmp >> 8; tmp |= tmp >> 16;
  tmp = (tmp + 1) >> 1;                                   /* floor power of two */
  if ((tmp << 1) < queue_length && tmp * thr_n != queue_length) tmp <<= 1;
  return tmp < max_bunch_size ? tmp : max_bunch_size;
}
uint32_t ffo_compute_job_number(uint32_t queue_length, uint32_t threads, uint32_t total_threads,
                                uint32_t max_bunch_size) {
  uint32_t left = queue_length;
  for (uint32_t i = 0; i != threads; ++i) {
    left -= ffo_compute_bunch_size(left, total_threads, max_bunch_size);
    if (left == 0) break;
  }
  return queue_length - left;
}

/* ------------------------------------------------------------------------------------------ */
/* small containers                                                                            */
/* ------------------------------------------------------------------------------------------ */
typedef struct { char** v; size_t n, cap; } strvec;
static void sv_push(strvec* s, const char* str, size_t len) {
  if (s->n == s->cap) {
    s->cap = s->cap ? 2 * s->cap : 16;
    s->v = (char**)realloc(s->v, s->cap * sizeof(char*));
  }
  char* c = (char*)malloc(len + 1);
  memcpy(c, str, len);
  c[len] = 0;
  s->v[s->n++] = c;
}
static void sv_free(strvec* s) {
  for (size_t i = 0; i < s->n; i++) free(s->v[i]);
  free(s->v);
  s->v = NULL;
  s->n = s->cap = 0;
}
typedef struct { char* v; size_t n, cap; } cbuf;
static void cb_push(cbuf* b, char c) {
  if (b->n + 1 >= b->cap) {
    b->cap = b->cap ? 2 * b->cap : 32;
    b->v = (char*)realloc(b->v, b->cap);
  }
  b->v[b->n++] = c;
  b->v[b->n] = 0;
}
static void cb_insert_front(cbuf* b, char c) {
  cb_push(b, 0);
  memmove(b->v + 1, b->v, b->n - 1);
  b->v[0] = c;
  b->v[b->n] = 0;
}
static void cb_clear(cbuf* b) {
  b->n = 0;
  if (b->v) b->v[0] = 0;
}

typedef struct { uint8_t op; uint64_t val; } token;
typedef struct {
  strvec rpn;
  token* tok; size_t n_tok;
} function;

struct ffo_parser {
  strvec vars;
  int check_is_equal;
  function* fun; size_t n_fun, cap_fun;
  uint64_t* eval_pos; size_t n_pos, cap_pos;
  uint64_t prime;
  /* duplicate detection */
  uint64_t *chk1, *chk2, *key1, *key2;
};

static long var_index(const ffo_parser* P, const char* name) {
  for (size_t i = 0; i < P->vars.n; i++)
    if (strcmp(P->vars.v[i], name) == 0) return (long)i;
  return -1;
}

/* ------------------------------------------------------------------------------------------ */
/* validate (ShuntingYardParser.cpp:778-837)                                                    */
/* ------------------------------------------------------------------------------------------ */
static char* validate(const char* in, uint64_t exp_n) {
  size_t len = strlen(in), size = len + 1;
  char* r = (char*)malloc(len + 3);
  memcpy(r, in, len + 1);
  r[len + 1] = 0;
  long* st = (long*)malloc((len + 1) * sizeof(long));
  size_t sp = 0;
  size_t i = 0;
  while (i != size) {
    if (r[i] == '+' && r[i + 1] == '-') r[i] = '$';
    if (r[i] == '-' && r[i + 1] == '+') r[i + 1] = '$';
    if (r[i] == '(') {
      st[sp++] = (i != 0 && r[i - 1] == '(') ? -(long)i : (long)i;
      i++;
    } else if (r[i] != ')') {
      i++;
    } else {
      if (sp == 0) {
        char num[32];
        snprintf(num, sizeof num, "%llu", (unsigned long long)exp_n);
        set_err("Mismatch of closing parentheses in expression ", num, ".");
        free(st); free(r);
        return NULL;
      }
      long top = st[sp - 1];
      if (i != size - 1 && r[i + 1] == ')' && top < 0) {
        r[-top] = '$';
        r[i] = '$';
      }
      sp--;
      i++;
    }
  }
  free(st);
  if (sp > 0) {
    char num[32];
    snprintf(num, sizeof num, "%llu", (unsigned long long)exp_n);
    set_err("Mismatch of opening parentheses in expression ", num, ".");
    free(r);
    return NULL;
  }
  size_t w = 0;
  for (size_t k = 0; k < len; k++)
    if (r[k] != '$') r[w++] = r[k];
  r[w] = 0;
  return r;
}

/* ------------------------------------------------------------------------------------------ */
/* parse (ShuntingYardParser.cpp:201-313); operator classes :484-524                            */
/* ------------------------------------------------------------------------------------------ */
static int get_weight(char c) {
  switch (c) {
    case '^': case '!': case '~': case ';': return 3;
    case '/': case '*': return 2;
    case '+': case '-': return 1;
    default: return 0;
  }
}
static int is_operator(char c) { return c == '+' || c == '-' || c == '*' || c == '/' || c == '^'; }
static int is_variable(char c) { return (c >= 'a' && c <= 'z') || (c >= 'A' && c <= 'Z'); }
static int is_operand(char c) { return !is_operator(c) && c != '(' && c != ')' && !is_variable(c) && c != ' '; }

static int parse(const char* validated, strvec* tokens) {
  /* global sign in front of a parenthesis (:204-210) */
  size_t len0 = strlen(validated);
  char* fun = (char*)malloc(len0 + 4);
  if (len0 > 2 && (validated[0] == '+' || validated[0] == '-') && validated[1] == '(') {
    if (validated[0] == '+') strcpy(fun, validated + 1);
    else { fun[0] = '0'; strcpy(fun + 1, validated); }
  } else {
    strcpy(fun, validated);
  }
  size_t len = strlen(fun);
  cbuf tmp = {0};
  char* op_stack = (char*)malloc(len + 1);
  size_t sp = 0;
  size_t c_counter = 0;
  int neg_exp = 0;
  int rc = 0;
  for (size_t k = 0; k < len; k++) {
    char ex = fun[k];
    char prev = c_counter ? fun[c_counter - 1] : 0;
    char next1 = c_counter + 1 <= len ? fun[c_counter + 1] : 0;
    char next2 = c_counter + 2 <= len ? fun[c_counter + 2] : 0;
    if (is_operand(ex) || is_variable(ex)) {
      cb_push(&tmp, ex);
    } else if (is_operator(ex)) {
      if (!neg_exp && tmp.n > 0) {
        sv_push(tokens, tmp.v, tmp.n);
        cb_clear(&tmp);
      }
      if (sp && prev == '(') {                      /* cases like +(-(x+...)) (:233-242) */
        if (next1 == '(') {
          sv_push(tokens, "0", 1);
          op_stack[sp++] = ex;
        } else if (!neg_exp) {
          cb_insert_front(&tmp, ex);
        } else {
          neg_exp = 0;
        }
      } else if (!sp && tokens->n == 0) {           /* leading sign (:243-244) */
        cb_insert_front(&tmp, ex);
      } else {
        while (sp && op_stack[sp - 1] != '(' && get_weight(op_stack[sp - 1]) >= get_weight(ex)) {
          sv_push(tokens, &op_stack[sp - 1], 1);
          sp--;
        }
        if (ex == '^' && tokens->n == 0) {
          set_err("malformed expression: '^' without a base", NULL, NULL);
          rc = -1;
          break;
        }
        char* back = ex == '^' ? tokens->v[tokens->n - 1] : NULL;
        int signed_back = back && strlen(back) > 1 && back[0] == '-' && prev != ')';
        if (ex == '^' && next1 == '(' && next2 == '-') {
          neg_exp = 1;
          if (signed_back) {
            memmove(back, back + 1, strlen(back));
            op_stack[sp++] = ';';
          } else {
            op_stack[sp++] = '~';
          }
        } else if (ex == '^' && signed_back) {
          op_stack[sp++] = '!';
          memmove(back, back + 1, strlen(back));
        } else {
          op_stack[sp++] = ex;
        }
      }
    } else if (ex == '(') {
      op_stack[sp++] = ex;
    } else if (ex == ')') {
      if (tmp.n > 0) sv_push(tokens, tmp.v, tmp.n);
      cb_clear(&tmp);
      while (sp) {
        char t = op_stack[--sp];
        if (t == '(') break;
        sv_push(tokens, &t, 1);
      }
    }
    ++c_counter;
  }
  if (rc == 0) {
    if (tmp.n > 0) sv_push(tokens, tmp.v, tmp.n);
    while (sp) sv_push(tokens, &op_stack[--sp], 1);
  }
  free(tmp.v);
  free(op_stack);
  free(fun);
  return rc;
}

/* ------------------------------------------------------------------------------------------ */
/* literals: precompute's operand branch (:668-711), FFInt(mpz_class) (FFInt.cpp:30-36)         */
/* ------------------------------------------------------------------------------------------ */
/* mpz_class(std::string) reads with base 0: 0x / 0b / leading-0 octal prefixes, optional '-' */
static int big_number(const char* s, uint64_t p, uint64_t* out) {
  int neg = 0;
  while (isspace((unsigned char)*s)) s++;
  if (*s == '-') { neg = 1; s++; }
  unsigned base = 10;
  if (s[0] == '0' && (s[1] == 'x' || s[1] == 'X')) { base = 16; s += 2; }
  else if (s[0] == '0' && (s[1] == 'b' || s[1] == 'B')) { base = 2; s += 2; }
  else if (s[0] == '0') base = 8;
  if (!*s) return -1;
  uint64_t r = 0;
  int any = 0;
  for (; *s; s++) {
    unsigned char ch = (unsigned char)*s;
    if (isspace(ch)) continue;
    unsigned d;
    if (ch >= '0' && ch <= '9') d = ch - '0';
    else if (ch >= 'a' && ch <= 'z') d = ch - 'a' + 10;
    else if (ch >= 'A' && ch <= 'Z') d = ch - 'A' + 10;
    else return -1;
    if (d >= base) return -1;
    r = (uint64_t)(((u128)r * base + d) % p);
    any = 1;
  }
  if (!any) return -1;
  *out = neg ? ffo_neg(r, p) : r;      /* in % p, then + p if negative */
  return 0;
}
/* std::stoull on a digit-initial string of at most 18 characters */
static uint64_t stoull_prefix(const char* s) {
  uint64_t v = 0;
  while (*s >= '0' && *s <= '9') v = v * 10 + (uint64_t)(*s++ - '0');
  return v;
}
/* FFInt(uint64) (FFInt.hpp:129-140) */
static uint64_t ffint(uint64_t n, uint64_t p) { return n < p ? n : n % p; }

/* classify one operand token: returns FFO_VARIABLE / FFO_NEG_VARIABLE / FFO_NUMBER or -1 */
static int operand_token(const ffo_parser* P, const char* token, uint64_t p, uint64_t* val) {
  size_t len = strlen(token);
  if (len > 18) {
    const char* t = token[0] == '+' ? token + 1 : token;
    if (big_number(t, p, val)) {
      set_err("cannot read integer '", token, "'");
      return -1;
    }
    return FFO_NUMBER;
  }
  int sign = token[0] == '-' ? -1 : token[0] == '+' ? 1 : 0;
  const char* name = sign ? token + 1 : token;
  long vi = var_index(P, name);
  if (vi >= 0) {
    *val = (uint64_t)vi;
    return sign < 0 ? FFO_NEG_VARIABLE : FFO_VARIABLE;
  }
  if (isdigit((unsigned char)name[0])) {
    uint64_t v = ffint(stoull_prefix(name), p);
    *val = sign < 0 ? ffo_neg(v, p) : v;
    return FFO_NUMBER;
  }
  set_err("Variable '", name, "' not declared!");
  return -1;
}

/* ------------------------------------------------------------------------------------------ */
/* the string-token evaluator (ShuntingYardParser.cpp:379-479), needed for duplicate detection */
/* and for re-evaluating a NUMBER's recipe after a prime change (:584-590)                     */
/* ------------------------------------------------------------------------------------------ */
static int evaluate_strings(const ffo_parser* P, const strvec* tokens, const uint64_t* values, uint64_t p,
                            uint64_t* result) {
  uint64_t* st = (uint64_t*)malloc((tokens->n + 1) * sizeof(uint64_t));
  size_t sp = 0;
  int rc = 0;
  for (size_t i = 0; i < tokens->n && rc == 0; i++) {
    const char* t = tokens->v[i];
    /* the operator test of THIS overload (:385) knows '~' and ';'; the public template
     * evaluate<FFIntTemp>(values) (ShuntingYardParser.hpp:249) does not, but nothing on this path calls it */
    if (t[1] == 0 && (t[0] == '+' || t[0] == '-' || t[0] == '*' || t[0] == '/' || t[0] == '^' || t[0] == '!' ||
                      t[0] == '~' || t[0] == ';')) {
      if (sp < 2) { set_err("malformed expression: operator '", t, "' lacks operands"); rc = -1; break; }
      uint64_t a = st[--sp];
      uint64_t* top = &st[sp - 1];
      switch (t[0]) {
        case '+': *top = ffo_add(*top, a, p); break;
        case '-': *top = ffo_sub(*top, a, p); break;
        case '*': *top = ffo_mul(*top, a, p); break;
        case '/': *top = ffo_div(*top, a, p); break;
        case '^': *top = ffo_pow(*top, a, p); break;
        case '!': *top = ffo_neg(ffo_pow(*top, a, p), p); break;
        case '~': *top = ffo_pow_int(*top, ffo_to_neg_int(a), p); break;                  /* :421-424 */
        case ';': *top = ffo_neg(ffo_pow_int(*top, ffo_to_neg_int(a), p), p); break;      /* :426-429 */
      }
    } else {
      uint64_t v;
      int k = operand_token(P, t, p, &v);
      if (k < 0) { rc = -1; break; }
      if (k == FFO_VARIABLE) st[sp++] = values ? values[v] : 0;
      else if (k == FFO_NEG_VARIABLE) st[sp++] = ffo_neg(values ? values[v] : 0, p);
      else st[sp++] = v;
    }
  }
  if (rc == 0) {
    if (sp) *result = st[sp - 1];
    else { set_err("Error in functional evaluation! Please check your input.", NULL, NULL); rc = -1; }
  }
  free(st);
  return rc;
}

/* ------------------------------------------------------------------------------------------ */
/* precompute (ShuntingYardParser.cpp:593-718)                                                  */
/* ------------------------------------------------------------------------------------------ */
static void free_tokens(function* f) {
  free(f->tok);
  f->tok = NULL;
  f->n_tok = 0;
}

static int precompute(const ffo_parser* P, function* f, uint64_t p) {
  free_tokens(f);
  size_t n = f->rpn.n;
  f->tok = (token*)calloc(n ? n : 1, sizeof(token));
  size_t j = 0;        /* emitted tokens */
  for (size_t k = 0; k < n; k++) {
    const char* t = f->rpn.v[k];
    if (t[1] == 0 && (is_operator(t[0]) || t[0] == '!' || t[0] == '~' || t[0] == ';')) {
      switch (t[0]) {
        case '+': f->tok[j++] = (token){FFO_PLUS, 0}; break;
        case '-': f->tok[j++] = (token){FFO_MINUS, 0}; break;
        case '*': f->tok[j++] = (token){FFO_MULT, 0}; break;
        case '^': f->tok[j++] = (token){FFO_POW, 0}; break;
        case '~': f->tok[j++] = (token){FFO_NEG_POW, 0}; break;
        case ';': f->tok[j++] = (token){FFO_NEG_POW_NEG, 0}; break;
        case '!': f->tok[j++] = (token){FFO_POW_NEG, 0}; break;
        case '/':
          if (j < 2) {      /* the reference indexes [j-1], [j-2] unconditionally: undefined there */
            set_err("malformed expression: operator '/' lacks operands", NULL, NULL);
            return -1;
          }
          if (f->tok[j - 1].op == FFO_NUMBER && f->tok[j - 2].op == FFO_NUMBER) {
            /* a rational coefficient: divide once (:621-639) */
            f->tok[j - 2].val = ffo_div(f->tok[j - 2].val, f->tok[j - 1].val, p);
            j -= 1;
          } else if (f->tok[j - 1].op == FFO_NUMBER &&
                     (f->tok[j - 2].op == FFO_VARIABLE || f->tok[j - 2].op == FFO_NEG_VARIABLE)) {
            /* x / c  ->  x * (1/c)  (:640-645) */
            f->tok[j - 1].val = ffo_div(1 % p, f->tok[j - 1].val, p);
            f->tok[j++] = (token){FFO_MULT, 0};
          } else {
            f->tok[j++] = (token){FFO_DIV, 0};
          }
          break;
      }
    } else {
      uint64_t v;
      int kind = operand_token(P, t, p, &v);
      if (kind < 0) return -1;
      f->tok[j++] = (token){(uint8_t)kind, v};
    }
  }
  f->n_tok = j;
  return 0;
}

/* ------------------------------------------------------------------------------------------ */
/* collection                                                                                  */
/* ------------------------------------------------------------------------------------------ */
static uint64_t splitmix(uint64_t* s) {
  uint64_t z = (*s += 0x9E3779B97F4A7C15ULL);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
  return z ^ (z >> 31);
}

static int add_function(ffo_parser* P, const char* text, uint64_t index) {
  char* v = validate(text, index);
  if (!v) return -1;
  function f;
  memset(&f, 0, sizeof f);
  if (parse(v, &f.rpn)) {
    free(v);
    sv_free(&f.rpn);
    return -1;
  }
  free(v);
  if (P->check_is_equal) {
    /* one random point in each of two primes (:125-140) */
    uint64_t p1 = ffo_prime(1), p0 = ffo_prime(0), v1, v2;
    if (evaluate_strings(P, &f.rpn, P->chk1, p1, &v1) || evaluate_strings(P, &f.rpn, P->chk2, p0, &v2)) {
      sv_free(&f.rpn);
      return -1;
    }
    for (size_t i = 0; i < P->n_fun; i++) {
      if (P->key1[i] == v1 && P->key2[i] == v2) {
        sv_free(&f.rpn);
        if (P->n_pos == P->cap_pos) {
          P->cap_pos = P->cap_pos ? 2 * P->cap_pos : 16;
          P->eval_pos = (uint64_t*)realloc(P->eval_pos, P->cap_pos * sizeof(uint64_t));
        }
        P->eval_pos[P->n_pos++] = i;
        return 0;
      }
    }
    if (P->n_pos == P->cap_pos) {
      P->cap_pos = P->cap_pos ? 2 * P->cap_pos : 16;
      P->eval_pos = (uint64_t*)realloc(P->eval_pos, P->cap_pos * sizeof(uint64_t));
    }
    P->eval_pos[P->n_pos++] = P->n_fun;
    if (P->n_fun == P->cap_fun) {
      P->cap_fun = P->cap_fun ? 2 * P->cap_fun : 16;
      P->fun = (function*)realloc(P->fun, P->cap_fun * sizeof(function));
      P->key1 = (uint64_t*)realloc(P->key1, P->cap_fun * sizeof(uint64_t));
      P->key2 = (uint64_t*)realloc(P->key2, P->cap_fun * sizeof(uint64_t));
    }
    P->key1[P->n_fun] = v1;
    P->key2[P->n_fun] = v2;
    P->fun[P->n_fun++] = f;
    return 0;
  }
  if (P->n_fun == P->cap_fun) {
    P->cap_fun = P->cap_fun ? 2 * P->cap_fun : 16;
    P->fun = (function*)realloc(P->fun, P->cap_fun * sizeof(function));
  }
  P->fun[P->n_fun++] = f;
  return 0;
}

static ffo_parser* new_parser(const char* const* vars, size_t n_vars, int check_is_equal) {
  ffo_parser* P = (ffo_parser*)calloc(1, sizeof *P);
  for (size_t i = 0; i < n_vars; i++) sv_push(&P->vars, vars[i], strlen(vars[i]));
  P->check_is_equal = check_is_equal;
  if (check_is_equal) {
    uint64_t s = 0x1234567887654321ULL;
    P->chk1 = (uint64_t*)malloc((n_vars + 1) * sizeof(uint64_t));
    P->chk2 = (uint64_t*)malloc((n_vars + 1) * sizeof(uint64_t));
    for (size_t i = 0; i < n_vars; i++) P->chk1[i] = 1 + splitmix(&s) % (ffo_prime(1) - 1);
    for (size_t i = 0; i < n_vars; i++) P->chk2[i] = 1 + splitmix(&s) % (ffo_prime(0) - 1);
  }
  return P;
}

int ffo_set_prime(ffo_parser* P, uint64_t p) {
  /* precompute_tokens (:550-591).  After a prime change the reference re-evaluates only the
   * NUMBER payloads from their `partial_rpn` recipes (the literals and divisions folded into
   * them); a fresh precompute for the new prime yields the same tokens and is what is done here.
   * tests/test_oracle.py checks both the first-prime and the prime-changed route against the
   * reference.  (One reference defect is NOT reproduced: for `x/(a/b)` the recipe keeps only its
   * first literal, :641, so after a prime change the reference multiplies by 1/a instead of b/a.) */
  for (size_t i = 0; i < P->n_fun; i++)
    if (precompute(P, &P->fun[i], p)) return -1;
  P->prime = p;
  return 0;
}

ffo_parser* ffo_parse(const char* const* funs, size_t n_fun, const char* const* vars, size_t n_vars,
                      int check_is_equal, uint64_t p) {
  ffo_parser* P = new_parser(vars, n_vars, check_is_equal);
  for (size_t i = 0; i < n_fun; i++) {
    if (add_function(P, funs[i], i)) { ffo_free(P); return NULL; }
  }
  if (ffo_set_prime(P, p)) { ffo_free(P); return NULL; }
  return P;
}

ffo_parser* ffo_parse_text(const char* text, const char* const* vars, size_t n_vars, int check_is_equal, uint64_t p) {
  /* file mode (:98-146): split on ';', drop ' ', '\t', '\n', skip empty chunks */
  ffo_parser* P = new_parser(vars, n_vars, check_is_equal);
  size_t len = strlen(text);
  char* chunk = (char*)malloc(len + 1);
  uint64_t line_c = 1;
  size_t start = 0;
  while (start < len) {
    size_t end = start;
    while (end < len && text[end] != ';') end++;
    size_t w = 0;
    for (size_t k = start; k < end; k++)
      if (text[k] != ' ' && text[k] != '\t' && text[k] != '\n') chunk[w++] = text[k];
    chunk[w] = 0;
    if (w > 0 && add_function(P, chunk, line_c)) { free(chunk); ffo_free(P); return NULL; }
    line_c++;
    start = end + 1;
  }
  free(chunk);
  if (ffo_set_prime(P, p)) { ffo_free(P); return NULL; }
  return P;
}

void ffo_free(ffo_parser* P) {
  if (!P) return;
  for (size_t i = 0; i < P->n_fun; i++) {
    free_tokens(&P->fun[i]);
    sv_free(&P->fun[i].rpn);
  }
  free(P->fun);
  free(P->eval_pos);
  free(P->chk1); free(P->chk2); free(P->key1); free(P->key2);
  sv_free(&P->vars);
  free(P);
}

size_t ffo_n_functions(const ffo_parser* P) { return P->n_fun; }
size_t ffo_n_out(const ffo_parser* P) { return P->check_is_equal ? P->n_pos : P->n_fun; }
size_t ffo_n_vars(const ffo_parser* P) { return P->vars.n; }
size_t ffo_rpn(const ffo_parser* P, size_t i, char* buf, size_t cap) {
  if (i >= P->n_fun) return 0;
  size_t need = 0;
  const strvec* r = &P->fun[i].rpn;
  for (size_t k = 0; k < r->n; k++) need += strlen(r->v[k]) + (k ? 1 : 0);
  if (buf && cap) {
    size_t w = 0;
    for (size_t k = 0; k < r->n && w + 1 < cap; k++) {
      if (k && w + 1 < cap) buf[w++] = ' ';
      for (const char* c = r->v[k]; *c && w + 1 < cap; c++) buf[w++] = *c;
    }
    buf[w] = 0;
  }
  return need;
}
size_t ffo_n_tokens(const ffo_parser* P, size_t i) { return i < P->n_fun ? P->fun[i].n_tok : 0; }
int ffo_token_op(const ffo_parser* P, size_t i, size_t j) { return P->fun[i].tok[j].op; }
uint64_t ffo_token_val(const ffo_parser* P, size_t i, size_t j) { return P->fun[i].tok[j].val; }
uint64_t ffo_eval_position(const ffo_parser* P, size_t k) { return P->check_is_equal ? P->eval_pos[k] : k; }

/* ------------------------------------------------------------------------------------------ */
/* evaluate_pre<FFInt> (ShuntingYardParser.hpp:365-479), one point at a time                    */
/* ------------------------------------------------------------------------------------------ */
int ffo_eval(const ffo_parser* P, const uint64_t* values, size_t n_points, uint64_t* out) {
  const uint64_t p = P->prime;
  const size_t n_vars = P->vars.n, n_out = ffo_n_out(P);
  size_t max_tok = 1;
  for (size_t i = 0; i < P->n_fun; i++)
    if (P->fun[i].n_tok > max_tok) max_tok = P->fun[i].n_tok;
  uint64_t* st = (uint64_t*)malloc((max_tok + 1) * sizeof(uint64_t));
  uint64_t* x = (uint64_t*)malloc((n_vars + 1) * sizeof(uint64_t));
  uint64_t* nx = (uint64_t*)malloc((n_vars + 1) * sizeof(uint64_t));
  uint64_t* res = (uint64_t*)malloc((P->n_fun + 1) * sizeof(uint64_t));
  int rc = 0;
  for (size_t pt = 0; pt < n_points && rc == 0; pt++) {
    for (size_t v = 0; v < n_vars; v++) {
      x[v] = values[v * n_points + pt];
      nx[v] = ffo_neg(x[v], p);                       /* neg_values (:369-374) */
    }
    for (size_t i = 0; i < P->n_fun && rc == 0; i++) {
      const function* f = &P->fun[i];
      size_t sp = 0;
      for (size_t j = 0; j < f->n_tok; j++) {
        const token t = f->tok[j];
        if (t.op <= FFO_NEG_POW_NEG) {
          if (sp < 2) { set_err("malformed expression: operator lacks operands", NULL, NULL); rc = -1; break; }
          const uint64_t a = st[--sp];
          uint64_t* top = &st[sp - 1];
          switch (t.op) {
            case FFO_PLUS: *top = ffo_add(*top, a, p); break;
            case FFO_MINUS: *top = ffo_sub(*top, a, p); break;
            case FFO_MULT: *top = ffo_mul(*top, a, p); break;
            case FFO_DIV: *top = ffo_div(*top, a, p); break;
            case FFO_POW: *top = ffo_pow(*top, a, p); break;
            case FFO_NEG_POW: *top = ffo_pow_int(*top, ffo_to_neg_int(a), p); break;
            case FFO_POW_NEG: *top = ffo_neg(ffo_pow(*top, a, p), p); break;
            case FFO_NEG_POW_NEG: *top = ffo_neg(ffo_pow_int(*top, ffo_to_neg_int(a), p), p); break;
          }
        } else if (t.op == FFO_VARIABLE) {
          st[sp++] = x[t.val];
        } else if (t.op == FFO_NEG_VARIABLE) {
          st[sp++] = nx[t.val];
        } else {
          st[sp++] = t.val;
        }
      }
      if (rc) break;
      if (!sp) { set_err("Error in functional evaluation! Please check your input.", NULL, NULL); rc = -1; break; }
      res[i] = st[sp - 1];
    }
    if (rc) break;
    for (size_t k = 0; k < n_out; k++) out[k * n_points + pt] = res[ffo_eval_position(P, k)];   /* (:462-478) */
  }
  free(st); free(x); free(nx); free(res);
  return rc;
}
