// compiler.hpp -- host-side tape compiler: text -> RPN -> expression trees ->
// accumulator-machine tape (tape.hpp) + per-prime constants.
//
// Behavioural contract (what a user of the reference's ShuntingYardParser sees):
//   * the same input grammar, including its quirks: ';'-separated files with all
//     blanks removed (reference: source/ShuntingYardParser.cpp:98-146), "+-" / "-+"
//     folding and "((..))" collapsing (validate, :778-837), unary signs glued to the
//     following operand, the synthetic operators '!' (-x^e), '~' (x^(-e)),
//     ';' (-x^(-e)), and ALL binary operators left associative, '^' included
//     (parse, :201-313; get_weight, :484-503);
//   * the same reverse Polish token strings (get_rp_functions, :526-528) -- exposed
//     through Collection::rpn for parity tests;
//   * literal handling of precompute (:593-718): tokens longer than 18 characters are
//     big integers reduced mod p, shorter ones are variables or stoull() numbers with
//     an optional sign prefix; unknown names are an error;
//   * check_is_equal duplicate detection (:73-92,125-140) with evaluation_positions
//     expansion (ShuntingYardParser.hpp:462-478);
//   * the values: every function evaluates to exactly the residue evaluate_pre returns.
//
// What is different is everything after the RPN: instead of a flat token list for a
// stack machine we build expression trees, fold every constant sub-expression once per
// prime with the reference's FFInt semantics (ffhost.hpp) -- a superset of the
// reference's NUM NUM / and VAR NUM / folds (:619-645) -- hash-cons each tree into a DAG so
// that equal sub-expressions are evaluated once, order sub-trees to minimise spill-stack
// depth, emit accumulator code (tape.hpp), fuse frequent instruction runs, and check the
// result: every function's tape is interpreted on the host at one point and compared with
// the value of its expression tree before the compilation is accepted.
#pragma once
#include <cstdint>
#include <string>
#include <unordered_map>
#include <vector>

#include "tape.hpp"

namespace ffgpu {

enum NodeKind : uint8_t {
  N_VAR, N_NEGVAR, N_NUM,
  N_ADD, N_SUB, N_MUL, N_DIV,
  N_POW,        // '^'
  N_NEGPOW,     // '~'  x^(-e)
  N_POWNEG,     // '!'  -(x^e)
  N_NEGPOWNEG,  // ';'  -(x^(-e))
};

struct Node {
  NodeKind kind;
  bool is_const;       // no variable below this node
  uint8_t need;        // spill-stack slots needed to bring this node into ACC
  uint32_t var;        // N_VAR / N_NEGVAR
  int32_t l, r;        // children (indices into the function's node array)
  uint32_t lit;        // N_NUM: index into Function::literals
};

// a number token exactly as it appears in the RPN ("12", "-3", "+5", "12132132323213213212")
struct Literal {
  std::string text;
};

struct Function {
  std::vector<std::string> rpn;      // reference-compatible RPN tokens
  std::vector<Node> nodes;           // post order: children precede parents
  std::vector<Literal> literals;
  int32_t root = -1;
};

// where a per-prime constant has to be written into the tape
struct ConstPatch {
  uint32_t instr;      // index into Collection::tape
  uint32_t fn;         // function
  int32_t node;        // constant sub-tree root
  bool inverse;        // store 1/value (division by a constant becomes a multiplication)
  bool negate;         // store -value (subtraction of a constant becomes an addition)
};

// algorithmic work of one probe (all functions), counted on the reference's
// precomputed tokens as SURVEY.md section 8(d) defines it
struct WorkCount {
  uint64_t tokens = 0;      // precomp_tokens the reference's evaluate_pre would walk
  uint64_t mm = 0;          // modular multiplications (literal powers: floor(log2 e)+popcount(e)-1)
  uint64_t ma = 0;          // additions / subtractions / negations
  uint64_t inv = 0;         // inversions (DIV and negative powers)
  uint64_t max_rpn_depth = 0;
};

struct Collection {
  std::vector<std::string> vars;
  std::unordered_map<std::string, uint32_t> var_index;
  bool check_is_equal = false;
  std::vector<Function> functions;            // distinct functions
  std::vector<uint64_t> evaluation_positions; // out index -> function (only if check_is_equal)
  size_t n_out = 0;

  std::vector<Instr> tape;                    // prime independent structure (imm = 0)
  std::vector<ConstPatch> patches;
  uint32_t stack_depth = 0;                   // spill slots the tape needs
  uint32_t n_temps = 0;                       // temp slots (shared sub-expressions, co-evaluated sums)
  uint32_t n_slots = 0;                       // shared-memory slots per probe: max over functions of spill + temps
                                              // (spill slots count up from 0, temp slots down from n_slots - 1)
  bool enable_fuse = true;                    // fuse frequent instruction runs (TERM, ACCUM2, ADDVARS, MULVARS)
  bool enable_cse = true;                     // share structurally equal sub-expressions
  bool enable_cross = true;                   // ... and across functions: whole functions of equal structure, denominators
  bool enable_table = true;                   // ... and sub-expressions used by SEVERAL functions (powers of a variable, common
                                              // factors, numerators): computed once per probe into extra variable rows
  uint32_t table_row_limit = 0;               // most shared rows (0: automatic)
  uint32_t n_table_rows = 0;                  // shared rows in use: the tape addresses them as variables n_vars, n_vars + 1, ...
  std::vector<Collection> pre;                // the collections whose outputs are the shared rows, evaluated before this tape at
                                              // the same points: level 0 (sub-expressions used in several places) over the
                                              // variables, level 1 (partial sums of long sums) over the variables and level 0
  bool enable_split = true;                   // cut long sums into partial sums (level-1 rows): intra-function parallelism
  bool enable_horner = true;                  // sums of monomials (expanded polynomials) are factored into Horner nests
  uint32_t n_split_rows = 0;                  // partial sums among the rows
  uint32_t n_input_vars = 0;                  // a rows' collection: how many of `vars` are the caller's (0: all of them)
  // first row each collection of `pre` writes
  const std::vector<uint32_t>& pre_row_offset() const { return pre_row_offset_; }
  uint32_t cse_fallbacks = 0;                 // functions emitted without sharing after a failed self-check
  uint32_t n_structural_duplicates = 0;       // functions of equal structure: evaluated once, stored to all their outputs
  uint32_t n_shared_denominators = 0;         // functions that reuse the inverted denominator of the function before them
  int temp_budget = 0;                        // temp slots per function (0: automatic, see FnEmitter)
  std::vector<uint32_t> fn_start;             // tape offset of every distinct function IN EVALUATION ORDER, then the end
  // chunk tables: for C = 1, 2, 4, ... the C + 1 offsets that cut the tape into C runs of whole
  // functions with about equal instruction counts (flattened: table c starts at chunk_table_off[c])
  std::vector<uint32_t> chunk_table;
  std::vector<uint32_t> chunk_table_off;
  unsigned max_chunks = 1;
  uint32_t exact_chunks = 0;                  // chunks of the one-function-per-chunk table (0: there is none), and its
  uint32_t exact_table = 0;                   // index into chunk_table_off
  uint64_t prime = 0;                         // prime the imm fields are resolved for (0: none)
  WorkCount work;                             // algorithmic: counted on the reference's token list
  WorkCount executed;                         // what the emitted tape does per probe (tokens = instructions,
                                              // max_rpn_depth = variable row reads)

  // parse + build trees + emit the tape.  Returns false and sets `err` on malformed input
  // (where the reference prints ERROR_MSG and exits, or runs into undefined behaviour).
  bool compile_strings(const std::vector<std::string>& funs, const std::vector<std::string>& vars,
                       bool check_is_equal, std::string& err);
  bool compile_text(const std::string& file_text, const std::vector<std::string>& vars, bool check_is_equal,
                    std::string& err);
  // recompute every constant for prime p and patch the tape (what prime_changed() ->
  // precompute_tokens() does in the reference, ShuntingYardParser.cpp:550-591)
  bool resolve(uint64_t p, std::string& err);
  uint64_t rebuilds = 0;                      // how often the tape has been generated (resolve() regenerates it when functions
                                              // were added or trimmed, and when the prime crosses 2^21: see resolve())
  // divide output `first` by the function `second` (written over factor_vars, matched to vars by position) and
  // rebuild the tape; replaces factors set earlier.  The prime has to be set again afterwards.
  bool apply_factors(const std::vector<std::pair<uint32_t, std::string>>& factors,
                     const std::vector<std::string>& factor_vars, std::string& err);
  // --- changing a compiled collection (the rest of ShuntingYardParser's public interface) ---
  // The tape is not rebuilt by these calls: `stale` is set and the next resolve() rebuilds it first, which is
  // where the reference pays too (precompute_tokens(true) after add_otf, ShuntingYardParser.cpp:550-591).
  bool stale = false;
  // add_otf (:347-377; validate = true, dedupe = true) and parse_function (:315-329; dedupe = false): append one
  // function.  With check_is_equal and dedupe, a function whose fingerprint is known returns the position of its
  // twin and gets an output of its own only if !no_duplicates.  `position` indexes `functions`.
  bool append(std::string text, bool validate, bool dedupe, bool no_duplicates, uint64_t& position, std::string& err);
  // add_otf_precompute (:331-341): the function arrives as reverse Polish tokens
  bool append_rpn(const std::vector<std::string>& rpn, uint64_t& position, std::string& err);
  // trim (:720-776): keep the listed functions (outputs, with check_is_equal) in their old order;
  // new_pos[k] is the new position of keep[k], ~0 if keep[k] did not exist
  bool trim(const std::vector<uint64_t>& keep, std::vector<uint64_t>& new_pos, std::string& err);

 private:
  bool add_function(std::string text, uint64_t index, bool strip_blanks, std::string& err, bool validate = true,
                    bool dedupe = true, bool no_duplicates = false, uint64_t* position = nullptr);
  void init_check_points();
  bool build_tree(Function& f, std::string& err);
  static void compute_needs(Function& f);
  std::vector<Function> base_functions_;      // one per output, before any factor was applied
  // what the tape is generated from: `functions` with every occurrence of a shared row replaced by a variable leaf
  std::vector<Function> lowered_;
  std::vector<uint32_t> pre_row_offset_;
  bool small_prime_mode_ = false;             // the tape was generated for a prime below 2^21 (no integer-exponent passes)
  bool resolve_constants(uint64_t p, std::string& err);
  const Function& efn(size_t i) const { return lowered_.empty() ? functions[i] : lowered_[i]; }
  void build_table(const std::vector<int32_t>& dup_of);
  void emit_function(uint32_t fn, const std::vector<uint32_t>& out_slots, int den_mode);
  bool check_function(uint32_t fn, uint64_t p, const uint64_t* point, uint64_t* temps, bool* temp_set) const;
  std::vector<uint8_t> den_mode_;             // per function: 0 as written, 1 parks 1/D in temp slot 0, 2 uses it
  std::vector<uint32_t> order_, pos_of_;      // evaluation order (position -> function) and its inverse
  std::vector<uint8_t> no_cse_fn_;
  void count_work(const Function& f);
  void peephole();
  void fuse();
  bool finish(std::string& err);
  // fingerprints for duplicate detection
  std::vector<uint64_t> chk_pts1_, chk_pts2_;
  std::unordered_map<std::string, uint64_t> chk_map_;
};

// --- pieces exposed for tests -------------------------------------------------------------
// validate(): "+-"/"-+" folding, "((x))" collapsing, parenthesis balance
bool validate_expr(std::string& s, uint64_t index, std::string& err);
// parse(): infix -> reference RPN tokens
bool parse_expr(std::string& s, std::vector<std::string>& rpn, std::string& err);
// value of a number token modulo p with the reference's rules
bool literal_value(const std::string& token, uint64_t p, uint64_t& out, std::string& err);
// evaluate all constant nodes of f for prime p (val[i] valid where nodes[i].is_const) and, if
// `point` is non-null, every node at that point (host fingerprinting only)
void eval_nodes(const Function& f, uint64_t p, const uint64_t* point, std::vector<uint64_t>& val);

}  // namespace ffgpu
