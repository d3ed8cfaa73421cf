/* ffgpu.h -- C ABI of the B200 probe evaluator (libffgpu.so).
 *
 * The drop-in boundary for FireFly's black-box hot path.  Everything FireFly's
 * Reconstructor asks of a parser-backed black box goes through these calls:
 *
 *   reference interface (jklappert/FireFly 2.0.3)                     replaced by
 *   ---------------------------------------------------------------  -------------------------
 *   ShuntingYardParser(file | strings, vars, check_is_equal)         ffgpu_compile_file /
 *     source/include/firefly/ShuntingYardParser.hpp:44,52             ffgpu_compile / _text
 *   ShuntingYardParser::precompute_tokens()  (prime_changed())       ffgpu_set_prime
 *     source/ShuntingYardParser.cpp:550-591; BlackBoxBase.hpp:43-45
 *   ShuntingYardParser::evaluate_pre<FFInt | FFIntVec<N>>(values)    ffgpu_eval
 *     source/include/firefly/ShuntingYardParser.hpp:94-95,365-479
 *     via BlackBoxBase::eval, BlackBoxBase.hpp:35-38; call sites
 *     Reconstructor.hpp:3383,3440, MPIWorker.hpp:475
 *   ShuntingYardParser::get_rp_function(i)                           ffgpu_rpn
 *     source/ShuntingYardParser.cpp:534-536
 *   ShuntingYardParser::add_otf / parse_function                     ffgpu_add_function
 *     source/ShuntingYardParser.cpp:347-377, 315-329
 *   ShuntingYardParser::add_otf_precompute   :331-341                ffgpu_add_function_rpn
 *   ShuntingYardParser::trim                 :720-776                ffgpu_trim
 *   copying a parser (black boxes hold one by value)                 ffgpu_tape_clone
 *   parsed_factors[i].evaluate_pre(values)[0] divided out            ffgpu_set_factors
 *     Reconstructor.hpp:3392-3396, 3449-3457                           (the division joins the tape)
 *   MPIWorker VALUES / RESULT messages  MPIWorker.hpp:213-228,479-501  ffgpu_eval_wire
 *   requested_probes / computed_probes, get_job, get_probe           ffgpu_queue_*
 *     Reconstructor.hpp:245-246, 2962-3012, 3326-3477
 *   primes()   source/ReconstHelper.cpp:24-125                       ffgpu_prime / ffgpu_n_primes
 *   Reconstructor::queue_probes point construction                   ffgpu_eval_line
 *     source/include/firefly/Reconstructor.hpp:2946-2960               (x_v = a_v * t + s_v on the GPU)
 *
 * Data layout: flat uint64_t residues, the layout std::vector<FFIntVec<N>> already has in
 * memory ([variable][point], contiguous) and that MPIWorker ships over the wire
 * (MPIWorker.hpp:213-228,479-501).  values[v * n_points + j] is variable v of point j,
 * out[f * n_points + j] is function f at point j.
 *
 * Errors: no call throws or exits.  Failing calls return NULL / non-zero and leave a message
 * for ffgpu_last_error() (thread local).  Where the reference prints ERROR_MSG and calls
 * std::exit (unknown variable, unbalanced parentheses, empty function) compilation fails.
 *
 * Threads: ffgpu_eval is re-entrant; `thread_id` picks one of the per-thread stream/staging
 * slots (Reconstructor calls the black box from thr_n pool threads, Reconstructor.hpp:2981).
 * ffgpu_set_prime must not run concurrently with ffgpu_eval on the same tape -- the same
 * contract prime_changed() has (called only after tp.kill_all(), Reconstructor.hpp:2403,2484).
 *
 * There is no CPU evaluation path: every ffgpu_eval* call runs the sm_100a kernels or fails.
 */
#ifndef FFGPU_H
#define FFGPU_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct ffgpu_ctx ffgpu_ctx;     /* a set of GPUs + streams + staging */
typedef struct ffgpu_tape ffgpu_tape;   /* a compiled collection, resident on every GPU of its ctx */

/* flags for ffgpu_compile*: */
enum {
  FFGPU_CHECK_IS_EQUAL = 1,   /* detect duplicate functions (ShuntingYardParser's check_is_equal) */
  FFGPU_NO_CSE = 2,           /* do not share equal sub-expressions inside a function (for A/B tests) */
  FFGPU_NO_FUSE = 4,          /* do not fuse instruction runs into TERM / MULADD / ACCUM2 / ADDVARS / MULVARS (for A/B tests) */
  FFGPU_NO_CROSS = 16,        /* do not share ACROSS functions: structural duplicates, shared denominators, shared rows (for A/B tests) */
  FFGPU_NO_TABLE = 32,        /* no shared rows (sub-expressions of several functions evaluated once per probe; for A/B tests) */
  FFGPU_NO_SPLIT = 64,        /* long sums stay in one piece (no partial-sum rows; for A/B tests) */
  FFGPU_NO_HORNER = 128,      /* sums of monomials are evaluated as written (no Horner nests; for A/B tests) */
  FFGPU_TEMPS_SHIFT = 8       /* bits 8..11: temp slots per function for shared sub-expressions (0 = automatic) */
};

/* ---- context ------------------------------------------------------------------------------ */
/* device_ids == NULL or n_devices == 0: the calling thread's current CUDA device. */
ffgpu_ctx* ffgpu_create(const int* device_ids, int n_devices);
void ffgpu_destroy(ffgpu_ctx*);
int ffgpu_n_devices(const ffgpu_ctx*);
const char* ffgpu_last_error(void);

/* ---- compile ------------------------------------------------------------------------------ */
/* ctx may be NULL: the handle then supports inspection (rpn, tape, work, set_prime) but refuses to
 * evaluate -- used to test the tape compiler on machines without a GPU. */
ffgpu_tape* ffgpu_compile(ffgpu_ctx*, const char* const* functions, size_t n_functions,
                          const char* const* vars, size_t n_vars, unsigned flags);
/* ';'-separated text with blanks/tabs/newlines ignored, as ShuntingYardParser reads a file */
ffgpu_tape* ffgpu_compile_text(ffgpu_ctx*, const char* text, const char* const* vars, size_t n_vars, unsigned flags);
ffgpu_tape* ffgpu_compile_file(ffgpu_ctx*, const char* path, const char* const* vars, size_t n_vars, unsigned flags);
/* an independent copy (ShuntingYardParser is copied by value into black boxes, e.g. test/test_bunch.cpp:33-34):
 * same functions, same prime, device buffers of its own */
ffgpu_tape* ffgpu_tape_clone(const ffgpu_tape*);
void ffgpu_tape_destroy(ffgpu_tape*);

size_t ffgpu_n_vars(const ffgpu_tape*);
size_t ffgpu_n_functions(const ffgpu_tape*);   /* distinct functions */
size_t ffgpu_n_out(const ffgpu_tape*);         /* values per point that ffgpu_eval returns */
size_t ffgpu_n_instr(const ffgpu_tape*);       /* tape length (without the terminators) */
size_t ffgpu_stack_depth(const ffgpu_tape*);    /* spill slots per probe */
size_t ffgpu_n_temps(const ffgpu_tape*);        /* temp slots per probe (shared sub-expressions) */
/* functions that failed the compiler's translation self-check with sharing on and were emitted as plain trees
 * (0 unless the sharing pass has a bug; a function failing as a plain tree too makes ffgpu_compile fail) */
size_t ffgpu_cse_fallbacks(const ffgpu_tape*);
/* functions whose expression is structurally equal to an earlier one of the collection (x+y and y+x are equal):
 * evaluated once, the value stored to every output -- what check_is_equal (reference ShuntingYardParser.cpp:125-140)
 * finds for functions equal AS WRITTEN, without the flag; FFGPU_NO_CSE turns it off */
size_t ffgpu_n_structural_duplicates(const ffgpu_tape*);
/* functions N/D whose denominator has the structure of the denominator of the function before them: 1/D is
 * computed once per run of such functions and kept in a temp slot (no denominator code, no inversion for the rest) */
size_t ffgpu_n_shared_denominators(const ffgpu_tape*);
/* shared rows: values computed once per probe by small tapes of their own and read by the functions' tape as extra
 * variables n_vars, n_vars + 1, ... (the reference shares nothing between functions and walks every token of every
 * function on one thread, ShuntingYardParser.hpp:365-479).  Two levels:
 *   level 0  sub-expressions that occur in several places of the collection (powers of a variable, common factors,
 *            numerators of different functions), over the variables;
 *   level 1  partial sums of LONG sums (a function of 10^3..10^5 terms is cut into pieces of about 300 tree nodes):
 *            the pieces are evaluated by different thread blocks, so a single large function fills the chip even at
 *            the Reconstructor's bunch sizes; over the variables and the level-0 rows.
 * ffgpu_get_table_tape copies the tape of a level (its output r is row offset(level) + r); FFGPU_NO_TABLE turns both
 * levels off, FFGPU_NO_SPLIT the second.  Independently of the rows, sums of monomials (expanded polynomials) are
 * factored into Horner nests where that saves multiplications (FFGPU_NO_HORNER: as written). */
size_t ffgpu_n_table_rows(const ffgpu_tape*);                     /* all levels */
size_t ffgpu_n_table_levels(const ffgpu_tape*);
size_t ffgpu_n_table_rows_level(const ffgpu_tape*, size_t level);
size_t ffgpu_table_row_offset(const ffgpu_tape*, size_t level);   /* first row the level writes */
size_t ffgpu_n_split_rows(const ffgpu_tape*);                     /* partial sums among them */
size_t ffgpu_get_table_tape(const ffgpu_tape*, size_t level, void* buf, size_t cap);
size_t ffgpu_n_slots(const ffgpu_tape*);        /* shared-memory slots per probe the kernel reserves (spill + temps) */
/* RPN of distinct function i, tokens joined by ' '; returns the length needed (without NUL) */
size_t ffgpu_rpn(const ffgpu_tape*, size_t i, char* buf, size_t cap);
/* copy up to cap 16-byte instructions of the resolved tape (layout: firefly_b200/csrc/tape.hpp) */
size_t ffgpu_get_tape(const ffgpu_tape*, void* buf, size_t cap);
/* algorithmic work of one probe counted on the reference's token list:
 * w[0]=tokens w[1]=modular multiplications w[2]=add/sub/neg w[3]=inversions w[4]=max RPN stack depth */
void ffgpu_work(const ffgpu_tape*, uint64_t w[5]);

/* work the compiled tape executes per probe (after constant folding and sharing of equal sub-expressions):
 * w[0]=instructions w[1]=modular multiplications w[2]=add/sub/neg w[3]=inversions w[4]=variable row reads */
void ffgpu_tape_work(const ffgpu_tape*, uint64_t w[5]);

/* Scanned factors: Reconstructor divides output i of the black box by parsed_factors[i].evaluate_pre(values)[0]
 * after every evaluation (Reconstructor.hpp:3392-3396, 3449-3457; MPIWorker.hpp:486-495).  Here the division becomes
 * part of function out_index[k] on the device: out = f / factor with FFInt's a/0 = 0.  The factor expressions are
 * written over factor_vars (FireFly uses x1..xn, ParserUtils.cpp:182-185), matched to the collection's variables by
 * position.  Replaces factors set earlier (n_factors = 0 removes them); needs quiescence; ffgpu_set_prime must be
 * called again afterwards. */
int ffgpu_set_factors(ffgpu_tape*, const uint32_t* out_index, const char* const* factors, size_t n_factors,
                      const char* const* factor_vars, size_t n_factor_vars);

/* ---- changing a compiled collection (the rest of ShuntingYardParser's public interface) ------
 * Host-side calls; the tape is rebuilt by the next ffgpu_set_prime (where the reference pays as well:
 * precompute_tokens(true), ShuntingYardParser.cpp:550-591).  Until then ffgpu_eval* refuses.  Need quiescence. */
#define FFGPU_ADD_PLAIN 1u          /* parse_function (:315-329): no duplicate check; refused on a check_is_equal tape */
#define FFGPU_ADD_VALIDATE 2u       /* parse_function's validate_fun; add_otf always validates */
#define FFGPU_ADD_NO_DUPLICATES 4u  /* add_otf's no_duplicates: a known function gets no further output */
/* add_otf (:347-377) / parse_function: append one function over the tape's variables.  Returns its position among
 * the distinct functions (the position of its twin if check_is_equal finds one), -1 on error. */
long ffgpu_add_function(ffgpu_tape*, const char* function, unsigned flags);
/* add_otf_precompute (:331-341): the function as reverse Polish tokens (what get_rp_function returns) */
long ffgpu_add_function_rpn(ffgpu_tape*, const char* const* tokens, size_t n_tokens);
/* trim (:720-776): keep the listed functions -- OUTPUT positions on a check_is_equal tape -- in their old order.
 * new_positions[k] = where keep[k] is now, ~0 if it did not exist. */
int ffgpu_trim(ffgpu_tape*, const uint64_t* keep, size_t n_keep, uint64_t* new_positions);
/* 1 if functions changed since the tape was built (ffgpu_set_prime pending) */
int ffgpu_is_stale(const ffgpu_tape*);

/* ---- prime -------------------------------------------------------------------------------- */
/* p must be a prime of the form 2^63 - c, 0 < c < 2^15 (all 300 primes of FireFly's table are). */
int ffgpu_set_prime(ffgpu_tape*, uint64_t p);
uint64_t ffgpu_get_prime(const ffgpu_tape*);
uint32_t ffgpu_n_primes(void);          /* 300 */
uint64_t ffgpu_prime(uint32_t i);       /* the i-th largest prime below 2^63 */

/* ---- evaluate ----------------------------------------------------------------------------- */
/* HOST buffers.  values: [n_vars][n_points] residues in [0, p) (what FFInt holds), out: [n_out][n_points].  Points are sharded over the
 * GPUs of the ctx by contiguous index range, copies and kernels are pipelined in chunks. */
int ffgpu_eval(ffgpu_tape*, const uint64_t* values, size_t n_points, uint64_t* out, uint32_t thread_id);
/* Probes on a line, the way Reconstructor::queue_probes builds them (Reconstructor.hpp:2946-2960):
 * point j has coordinates x_v = a[v] * t[j] + s[v] mod p (a[0] = 1 and a[i] = rand_zi[i-1], s = shift for the
 * normal case; a[v] in {0,1}, s[v] = fixed value for the factor scan).  Only t crosses PCIe (8 bytes per probe
 * instead of 8 * n_vars); the coordinates are formed on the GPU.  a, s: [n_vars] canonical residues, t: [n_points],
 * out: [n_out][n_points], all HOST buffers.  Values are identical to ffgpu_eval on the expanded points. */
int ffgpu_eval_line(ffgpu_tape*, const uint64_t* a, const uint64_t* s, const uint64_t* t, size_t n_points,
                    uint64_t* out, uint32_t thread_id);
/* The MPI worker's message format (MPIWorker.hpp:213-228, 479-501): values_msg = [index, x_1..x_n] * k as the master
 * sends it with tag VALUES, result_msg = [index, f_1..f_m] * k as the worker answers with tag RESULT (the caller
 * appends its free-slot count).  One batch evaluation for all k probes. */
int ffgpu_eval_wire(ffgpu_tape*, const uint64_t* values_msg, size_t k, uint64_t* result_msg, uint32_t thread_id);
/* DEVICE buffers on GPU `device_index` of the ctx, row strides in elements; launches on `stream`
 * (a cudaStream_t, NULL = default stream) and returns without synchronising. */
int ffgpu_eval_device(ffgpu_tape*, int device_index, const uint64_t* d_values, size_t in_stride, size_t n_points,
                      uint64_t* d_out, size_t out_stride, void* stream);
/* pinned host memory for ffgpu_eval buffers (pageable memory works too, slower) */
void* ffgpu_host_alloc(size_t bytes);
void ffgpu_host_free(void*);

/* ---- probe queue ---------------------------------------------------------------------------
 * The pair of queues Reconstructor keeps around its black box -- requested_probes (index, values) filled by
 * queue_probes, computed_probes drained by get_probe (Reconstructor.hpp:245-246, 2962-2971, 2996-3012) -- with
 * ONE consumer in place of the thr_n get_job threads (:3326-3477): a worker thread that takes everything
 * waiting (up to max_batch probes, 0 = 65536) and evaluates it in one launch.  Where the reference's bunch
 * policy splits the backlog over threads in bunches of at most 128 (compute_bunch_size, utils.cpp:305-330), this
 * one lets batches grow with the backlog.  Every probe gets the next running index (Reconstructor's `ind`);
 * results come back tagged with it.  Thread safe.  Drain the queue before ffgpu_set_prime (as Reconstructor
 * does with tp.kill_all() before a prime change, :2403). */
typedef struct ffgpu_queue ffgpu_queue;
ffgpu_queue* ffgpu_queue_create(ffgpu_tape*, size_t max_batch);
void ffgpu_queue_destroy(ffgpu_queue*);
/* queue one point (values[n_vars], residues in [0,p)); returns its index */
uint64_t ffgpu_queue_push(ffgpu_queue*, const uint64_t* values);
/* queue n points, point-major values[n][n_vars] (the layout of requested_probes); returns the first index */
uint64_t ffgpu_queue_push_many(ffgpu_queue*, const uint64_t* values, size_t n);
/* queue n points on the line x_v = a[v]*t[j] + s[v] (what queue_probes generates, :2946-2960): only t is
 * shipped to the GPU; returns the first index */
uint64_t ffgpu_queue_push_line(ffgpu_queue*, const uint64_t* a, const uint64_t* s, const uint64_t* t, size_t n);
/* fetch up to cap finished probes: indices[i] and out[f * cap + i] (function-major, row stride cap).  wait != 0
 * blocks until at least one probe is ready or nothing is outstanding.  Returns the count, -1 after an
 * evaluation error (ffgpu_queue_error). */
long ffgpu_queue_pop(ffgpu_queue*, uint64_t* indices, uint64_t* out, size_t cap, int wait);
size_t ffgpu_queue_pending(ffgpu_queue*);                 /* queued + being evaluated */
void ffgpu_queue_stats(ffgpu_queue*, uint64_t s[3]);      /* launches' worth of batches, probes done, largest batch */
const char* ffgpu_queue_error(ffgpu_queue*);

/* ---- tuning / measurement ----------------------------------------------------------------- */
/* probes per thread (1, 2, 4, 8; 0 = automatic) and threads per block (128, 256, 512; 0 = automatic) */
int ffgpu_set_launch(ffgpu_tape*, int probes_per_thread, int threads_per_block);
/* Aggregator of concurrent small ffgpu_eval calls (a few hundred points at most per call, many caller threads -- the
 * Reconstructor's thread pool, Reconstructor.hpp:3325-3477 with bunches from utils.cpp:305-330): calls that arrive while
 * the batch before them is on the GPU share ONE launch; a lone caller is served at once.  enable: 1 on, 0 off (every
 * call its own launch on its caller slot's stream), -1 environment FFGPU_COALESCE / default OFF: measured under the
 * unmodified Reconstructor, sharing cuts the launches to an eighth but costs latency (callers wait for each other), and
 * the reference's own average black-box time is lower without it.  max_points: points per shared launch (0 =
 * automatic, 512..4096 by the collection's width) -- what GpuBlackBox::set_bunch_size steers.  The values are the same
 * either way.  Stats: launches shared so far, calls served by them, points. */
int ffgpu_set_coalesce(ffgpu_tape*, int enable, size_t max_points);
void ffgpu_coalesce_stats(const ffgpu_tape*, uint64_t stats[3]);
/* points per pipelined chunk of ffgpu_eval (0 = automatic) */
int ffgpu_set_chunk(ffgpu_tape*, size_t points);
/* kernels launched by this library in this process so far */
uint64_t ffgpu_launch_count(void);
/* synthetic probe points on the device: d_values[v*stride + j] = 1 + splitmix64(seed + v*total_points +
 * first_point + j) mod (p-1), j < n_points */
int ffgpu_fill_points(ffgpu_ctx*, int device_index, uint64_t* d_values, size_t n_vars, size_t n_points, size_t stride,
                      uint64_t first_point, uint64_t total_points, uint64_t seed, uint64_t p, void* stream);
/* d_sums[2f] = sum of row f mod p, d_sums[2f+1] = xor of row f */
int ffgpu_checksum(ffgpu_ctx*, int device_index, const uint64_t* d_out, size_t n_out, size_t n_points, size_t stride,
                   uint64_t p, uint64_t* d_sums, void* stream);
/* device field arithmetic self test (tests only): r[i] = op(a[i], b[i]), HOST buffers.
 * op: 0 add 1 sub 2 mul 3 neg 4 inverse(Fermat) 5 pow(residue exponent) 6 inverse(block batched)
 *     7 pow(-(int)b) 8 canonical 9 1/(a*b) through the batched inverse
 *     10 (a+b)*b 11 (a-b)*b 12 (a*b)^2-b 13 (-a)*b+a  (chains on un-canonicalised intermediates)
 * a, b: any representatives below 2^63 + 2^32 (the interpreter's bounded lazy residues). */
int ffgpu_field_ops(ffgpu_ctx*, int op, const uint64_t* a, const uint64_t* b, uint64_t* r, size_t n, uint64_t p);

#ifdef __cplusplus
}
#endif
#endif /* FFGPU_H */
