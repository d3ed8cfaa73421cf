/* ff_oracle.h -- TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * Plain-C restatement of the reference's CPU hot path (jklappert/FireFly 2.0.3):
 * text -> RPN (ShuntingYardParser::validate/parse), RPN -> precomputed tokens
 * (ShuntingYardParser::precompute) and the stack-machine evaluator
 * (ShuntingYardParser::evaluate_pre<FFInt>) over FFInt's DEFAULT arithmetic.
 * Every function in ff_oracle.c cites the reference file:line it follows.
 *
 * Pinned against: oracle/_ref (the unmodified reference compiled here by
 * oracle/Makefile) in tests/test_oracle_vs_ref.py, and against the committed
 * golden vectors in tests/golden/ generated from oracle/_ref by
 * tests/golden/make_golden.py.
 */
#ifndef FF_ORACLE_H
#define FF_ORACLE_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* opcode numbering of the reference (ShuntingYardParser.hpp:220-238) */
enum ffo_token {
  FFO_PLUS = 0, FFO_MINUS, FFO_MULT, FFO_DIV, FFO_POW, FFO_NEG_POW, FFO_POW_NEG, FFO_NEG_POW_NEG,
  FFO_OPERATOR, FFO_VARIABLE, FFO_NEG_VARIABLE, FFO_NUMBER
};

typedef struct ffo_parser ffo_parser;

/* field arithmetic mod p (FFInt.cpp:130-192, 226-228, 356-405; FFInt.hpp:129-174) */
uint64_t ffo_add(uint64_t a, uint64_t b, uint64_t p);
uint64_t ffo_sub(uint64_t a, uint64_t b, uint64_t p);
uint64_t ffo_neg(uint64_t a, uint64_t p);
uint64_t ffo_mul(uint64_t a, uint64_t b, uint64_t p);
uint64_t ffo_inv(uint64_t a, uint64_t p);
uint64_t ffo_div(uint64_t a, uint64_t b, uint64_t p);
uint64_t ffo_pow(uint64_t a, uint64_t e, uint64_t p);      /* FFInt::pow(const FFInt&) */
uint64_t ffo_pow_int(uint64_t a, int32_t e, uint64_t p);   /* FFInt::pow<int>(e)       */
int32_t ffo_to_neg_int(uint64_t a);                         /* FFInt::to_neg_int        */

/* parse a collection given as strings (ShuntingYardParser.cpp:45-58,147-178) or as the
 * text of a ';'-separated file (ShuntingYardParser.cpp:32-43,98-146).  Returns NULL on a
 * parse error (the reference prints ERROR_MSG and exits); ffo_last_error() says why. */
ffo_parser* ffo_parse(const char* const* funs, size_t n_fun, const char* const* vars, size_t n_vars,
                      int check_is_equal, uint64_t p);
ffo_parser* ffo_parse_text(const char* text, const char* const* vars, size_t n_vars, int check_is_equal, uint64_t p);
void ffo_free(ffo_parser*);
const char* ffo_last_error(void);

/* recompute the NUMBER payloads for a new prime (precompute_tokens after a prime change,
 * ShuntingYardParser.cpp:550-591; computed as a fresh precompute) */
int ffo_set_prime(ffo_parser*, uint64_t p);

size_t ffo_n_functions(const ffo_parser*);   /* distinct functions (after check_is_equal) */
size_t ffo_n_out(const ffo_parser*);         /* values returned by evaluate_pre */
size_t ffo_n_vars(const ffo_parser*);
/* RPN of function i joined with ' ' (get_rp_function); returns the full length */
size_t ffo_rpn(const ffo_parser*, size_t i, char* buf, size_t cap);
/* precomputed tokens of function i */
size_t ffo_n_tokens(const ffo_parser*, size_t i);
int ffo_token_op(const ffo_parser*, size_t i, size_t j);
uint64_t ffo_token_val(const ffo_parser*, size_t i, size_t j);
uint64_t ffo_eval_position(const ffo_parser*, size_t k);

/* evaluate_pre<FFInt> at n_points points.  values: [n_vars][n_points], out: [n_out][n_points].
 * Returns 0, or -1 if a function leaves an empty stack (reference: ERROR_MSG + exit,
 * ShuntingYardParser.hpp:454-459). */
int ffo_eval(const ffo_parser*, const uint64_t* values, size_t n_points, uint64_t* out);

/* the bunch policy of Reconstructor::get_job (utils.cpp:305-344) */
uint32_t ffo_compute_bunch_size(uint32_t queue_length, uint32_t thr_n, uint32_t max_bunch_size);
uint32_t ffo_compute_job_number(uint32_t queue_length, uint32_t threads, uint32_t total_threads, uint32_t max_bunch_size);

/* prime table (ReconstHelper.cpp:24-125): p_i = 2^63 - c_i */
uint32_t ffo_n_primes(void);
uint64_t ffo_prime(uint32_t i);

#ifdef __cplusplus
}
#endif
#endif
