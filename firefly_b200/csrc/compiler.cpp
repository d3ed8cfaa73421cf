// compiler.cpp -- see compiler.hpp.  Host-only C++ (no CUDA).
#include "compiler.hpp"

#include <algorithm>
#include <cctype>
#include <cstring>

#include "ffhost.hpp"

namespace ffgpu {

// ------------------------------------------------------------------------------------------
// 1. validate  (behaviour of reference ShuntingYardParser::validate, ShuntingYardParser.cpp:778-837)
// ------------------------------------------------------------------------------------------
bool validate_expr(std::string& s, uint64_t index, std::string& err) {
  // Work on the characters plus a terminating NUL; deleted characters are marked and
  // squeezed out at the end.  The scan is strictly left to right and looks at the live
  // buffer, so marks made earlier are visible to later tests (as in the reference).
  const char DEL = '$';
  std::vector<char> b(s.begin(), s.end());
  b.push_back('\0');
  const size_t n = b.size();
  std::vector<long> open;               // positions of '(' ; negative = directly preceded by '('
  for (size_t i = 0; i < n; ++i) {
    const char ch = b[i];
    if (ch == '+' && b[i + 1] == '-') b[i] = DEL;          // "+-" -> "-"
    if (b[i] == '-' && b[i + 1] == '+') b[i + 1] = DEL;    // "-+" -> "-"
    if (b[i] == '(') {
      open.push_back((i != 0 && b[i - 1] == '(') ? -(long)i : (long)i);
    } else if (b[i] == ')') {
      if (open.empty()) {
        err = "Mismatch of closing parentheses in expression " + std::to_string(index) + ".";
        return false;
      }
      const long top = open.back();
      if (i + 1 < n && b[i + 1] == ')' && top < 0) {        // "((x))" -> "(x)"
        b[(size_t)(-top)] = DEL;
        b[i] = DEL;
      }
      open.pop_back();
    }
  }
  if (!open.empty()) {
    err = "Mismatch of opening parentheses in expression " + std::to_string(index) + ".";
    return false;
  }
  std::string out;
  out.reserve(s.size());
  for (size_t i = 0; i + 1 < n; ++i)
    if (b[i] != DEL) out.push_back(b[i]);
  s.swap(out);
  return true;
}

// ------------------------------------------------------------------------------------------
// 2. parse: infix -> RPN  (behaviour of reference ShuntingYardParser::parse, :201-313)
// ------------------------------------------------------------------------------------------
namespace {
inline bool is_op_char(char c) { return c == '+' || c == '-' || c == '*' || c == '/' || c == '^'; }
inline bool is_letter(char c) { return (c >= 'a' && c <= 'z') || (c >= 'A' && c <= 'Z'); }
inline int weight(char c) {
  switch (c) {
    case '^': case '!': case '~': case ';': return 3;
    case '*': case '/': return 2;
    case '+': case '-': return 1;
    default: return 0;
  }
}
}  // namespace

bool parse_expr(std::string& s, std::vector<std::string>& out, std::string& err) {
  // a global sign in front of a parenthesis: "+(" loses the sign, "-(" becomes "0-("
  if (s.size() > 2 && (s[0] == '+' || s[0] == '-') && s[1] == '(') {
    if (s[0] == '+') s.erase(0, 1);
    else s.insert(s.begin(), '0');
  }
  out.clear();
  std::string cur;                  // operand being collected
  std::vector<char> ops;            // operator stack
  bool neg_exp = false;             // inside "^(-": the '-' is still to be swallowed
  const size_t n = s.size();
  auto at = [&](size_t i) -> char { return i < n ? s[i] : '\0'; };
  size_t pos = 0;                   // counts non-NUL characters like the reference's c_counter
  for (size_t idx = 0; idx < n; ++idx) {
    const char ch = s[idx];
    if (ch == '\0') continue;
    if (ch == '(') {
      ops.push_back(ch);
    } else if (ch == ')') {
      if (!cur.empty()) out.push_back(cur);
      cur.clear();
      while (!ops.empty()) {
        const char t = ops.back();
        ops.pop_back();
        if (t == '(') break;
        out.emplace_back(1, t);
      }
    } else if (is_op_char(ch)) {
      if (!neg_exp && !cur.empty()) {
        out.push_back(cur);
        cur.clear();
      }
      if (!ops.empty() && pos > 0 && at(pos - 1) == '(') {
        // a sign directly behind an opening parenthesis
        if (at(pos + 1) == '(') {           // "(-(" : 0 - (...)
          out.emplace_back("0");
          ops.push_back(ch);
        } else if (!neg_exp) {
          cur.insert(cur.begin(), ch);      // glue the sign to the operand that follows
        } else {
          neg_exp = false;                  // the '-' of "^(-"
        }
      } else if (ops.empty() && out.empty()) {
        cur.insert(cur.begin(), ch);        // sign at the very beginning
      } else {
        while (!ops.empty() && ops.back() != '(' && weight(ops.back()) >= weight(ch)) {
          out.emplace_back(1, ops.back());
          ops.pop_back();
        }
        if (ch == '^') {
          if (out.empty()) {
            err = "malformed expression: '^' without a base";
            return false;
          }
          std::string& base = out.back();
          const bool signed_base = base.size() > 1 && base.front() == '-' && at(pos - 1) != ')';
          if (at(pos + 1) == '(' && at(pos + 2) == '-') {
            neg_exp = true;
            if (signed_base) { base.erase(base.begin()); ops.push_back(';'); }
            else ops.push_back('~');
          } else if (signed_base) {
            base.erase(base.begin());
            ops.push_back('!');
          } else {
            ops.push_back('^');
          }
        } else {
          ops.push_back(ch);
        }
      }
    } else if (ch != ' ') {
      cur.push_back(ch);                    // digits, letters and anything else
    }
    ++pos;
  }
  if (!cur.empty()) out.push_back(cur);
  while (!ops.empty()) {
    out.emplace_back(1, ops.back());
    ops.pop_back();
  }
  return true;
}

// ------------------------------------------------------------------------------------------
// 3. literals  (behaviour of precompute's operand branch, :668-711, and FFInt(mpz_class), FFInt.cpp:30-36)
// ------------------------------------------------------------------------------------------
namespace {
// GMP's mpz_set_str(.., base 0) as reached through mpz_class(std::string): optional '-',
// then 0x/0X hex, 0b/0B binary, leading 0 octal, otherwise decimal.
bool big_literal(const std::string& t, uint64_t p, uint64_t& out) {
  size_t i = 0;
  bool negative = false;
  while (i < t.size() && std::isspace((unsigned char)t[i])) ++i;
  if (i < t.size() && t[i] == '-') { negative = true; ++i; }
  unsigned base = 10;
  if (i + 1 < t.size() && t[i] == '0' && (t[i + 1] == 'x' || t[i + 1] == 'X')) { base = 16; i += 2; }
  else if (i + 1 < t.size() && t[i] == '0' && (t[i + 1] == 'b' || t[i + 1] == 'B')) { base = 2; i += 2; }
  else if (i < t.size() && t[i] == '0') { base = 8; }
  if (i >= t.size()) return false;
  uint64_t r = 0;
  bool any = false;
  for (; i < t.size(); ++i) {
    const unsigned char ch = (unsigned char)t[i];
    if (std::isspace(ch)) continue;
    unsigned d;
    if (ch >= '0' && ch <= '9') d = ch - '0';
    else if (ch >= 'a' && ch <= 'z') d = ch - 'a' + 10;
    else if (ch >= 'A' && ch <= 'Z') d = ch - 'A' + 10;
    else return false;
    if (d >= base) return false;
    r = (uint64_t)(((ffhost::u128)r * base + d) % p);
    any = true;
  }
  if (!any) return false;
  out = negative ? ffhost::neg(r, p) : r;
  return true;
}
// std::stoull on a digit-initial string: the leading run of digits (<= 18 characters: no overflow)
uint64_t leading_digits(const std::string& t, size_t from) {
  uint64_t v = 0;
  for (size_t i = from; i < t.size() && t[i] >= '0' && t[i] <= '9'; ++i) v = v * 10 + (uint64_t)(t[i] - '0');
  return v;
}
}  // namespace

bool literal_value(const std::string& token, uint64_t p, uint64_t& out, std::string& err) {
  if (token.size() > 18) {
    const std::string t = token[0] == '+' ? token.substr(1) : token;
    if (!big_literal(t, p, out)) {
      err = "cannot read integer '" + token + "'";
      return false;
    }
    return true;
  }
  size_t from = 0;
  bool negative = false;
  if (!token.empty() && (token[0] == '-' || token[0] == '+')) { negative = token[0] == '-'; from = 1; }
  const uint64_t v = leading_digits(token, from) % p;
  out = negative ? ffhost::neg(v, p) : v;
  return true;
}

// ------------------------------------------------------------------------------------------
// 4. RPN -> expression tree
// ------------------------------------------------------------------------------------------
bool Collection::build_tree(Function& f, std::string& err) {
  f.nodes.clear();
  f.literals.clear();
  f.nodes.reserve(f.rpn.size());
  std::vector<int32_t> st;
  for (const std::string& tok : f.rpn) {
    Node nd{};
    nd.l = nd.r = -1;
    const bool is_operator = tok.size() == 1 && (is_op_char(tok[0]) || tok[0] == '!' || tok[0] == '~' || tok[0] == ';');
    if (is_operator) {
      if (st.size() < 2) {
        err = "malformed expression: operator '" + tok + "' lacks operands";
        return false;
      }
      nd.r = st.back(); st.pop_back();
      nd.l = st.back(); st.pop_back();
      switch (tok[0]) {
        case '+': nd.kind = N_ADD; break;
        case '-': nd.kind = N_SUB; break;
        case '*': nd.kind = N_MUL; break;
        case '/': nd.kind = N_DIV; break;
        case '^': nd.kind = N_POW; break;
        case '~': nd.kind = N_NEGPOW; break;
        case '!': nd.kind = N_POWNEG; break;
        default: nd.kind = N_NEGPOWNEG; break;
      }
      nd.is_const = f.nodes[nd.l].is_const && f.nodes[nd.r].is_const;
    } else {
      // operand: big integer, signed/unsigned variable, or number
      bool as_number = false;
      if (tok.size() > 18) {
        as_number = true;
      } else {
        const bool sgn = !tok.empty() && (tok[0] == '-' || tok[0] == '+');
        const std::string name = sgn ? tok.substr(1) : tok;
        auto it = var_index.find(name);
        if (it != var_index.end()) {
          nd.kind = (sgn && tok[0] == '-') ? N_NEGVAR : N_VAR;
          nd.var = it->second;
          nd.is_const = false;
        } else if (!name.empty() && std::isdigit((unsigned char)name[0])) {
          as_number = true;
        } else {
          err = "Variable '" + name + "' not declared!";
          return false;
        }
      }
      if (as_number) {
        uint64_t probe;
        if (!literal_value(tok, 9223372036854775783ULL, probe, err)) return false;
        nd.kind = N_NUM;
        nd.is_const = true;
        nd.lit = (uint32_t)f.literals.size();
        f.literals.push_back(Literal{tok});
      }
    }
    st.push_back((int32_t)f.nodes.size());
    f.nodes.push_back(nd);
  }
  if (st.empty()) {
    err = "Error in functional evaluation! Please check your input.";
    return false;
  }
  f.root = st.back();      // evaluate_pre returns the top of the stack

  compute_needs(f);
  return true;
}

// spill-stack need (Sethi-Ullman): operands that can be addressed directly cost nothing
void Collection::compute_needs(Function& f) {
  auto direct = [&](int32_t i) {
    const Node& n = f.nodes[i];
    return n.is_const || n.kind == N_VAR || n.kind == N_NEGVAR;
  };
  for (size_t i = 0; i < f.nodes.size(); ++i) {
    Node& n = f.nodes[i];
    if (direct((int32_t)i)) { n.need = 0; continue; }
    const bool dl = direct(n.l), dr = direct(n.r);
    // a constant exponent / divisor rides in the instruction
    if (dr) n.need = f.nodes[n.l].need;
    else if (dl) n.need = f.nodes[n.r].need;
    else {
      const unsigned a = f.nodes[n.l].need, b = f.nodes[n.r].need;
      const unsigned lf = std::max(a, b + 1), rf = std::max(b, a + 1);
      const unsigned m = std::min(lf, rf);
      n.need = (uint8_t)std::min(m, 255u);
    }
  }
}

// ------------------------------------------------------------------------------------------
// 5. node evaluation on the host (constants per prime; fingerprints)
// ------------------------------------------------------------------------------------------
void eval_nodes(const Function& f, uint64_t p, const uint64_t* point, std::vector<uint64_t>& val) {
  val.assign(f.nodes.size(), 0);
  std::string err;
  for (size_t i = 0; i < f.nodes.size(); ++i) {
    const Node& n = f.nodes[i];
    if (!n.is_const && !point) continue;
    uint64_t v = 0;
    switch (n.kind) {
      case N_VAR: v = point[n.var] % p; break;
      case N_NEGVAR: v = ffhost::neg(point[n.var] % p, p); break;
      case N_NUM: literal_value(f.literals[n.lit].text, p, v, err); break;
      case N_ADD: v = ffhost::add(val[n.l], val[n.r], p); break;
      case N_SUB: v = ffhost::sub(val[n.l], val[n.r], p); break;
      case N_MUL: v = ffhost::mul(val[n.l], val[n.r], p); break;
      case N_DIV: v = ffhost::div(val[n.l], val[n.r], p); break;
      case N_POW: v = ffhost::pow_ff(val[n.l], val[n.r], p); break;
      case N_POWNEG: v = ffhost::neg(ffhost::pow_ff(val[n.l], val[n.r], p), p); break;
      case N_NEGPOW: v = ffhost::pow_int(val[n.l], ffhost::to_neg_int(val[n.r]), p); break;
      case N_NEGPOWNEG: v = ffhost::neg(ffhost::pow_int(val[n.l], ffhost::to_neg_int(val[n.r]), p), p); break;
    }
    val[i] = v;
  }
}

// ------------------------------------------------------------------------------------------
// 6. algorithmic work on the reference's token list
// ------------------------------------------------------------------------------------------
void Collection::count_work(const Function& f) {
  // Replay precompute's folding on the RPN to know which tokens evaluate_pre would walk.
  enum K : uint8_t { OPR, VARI, NUMB };
  struct T { K k; char op; uint64_t lit; bool small; };
  std::vector<T> tk;
  for (const std::string& s : f.rpn) {
    const bool is_operator = s.size() == 1 && (is_op_char(s[0]) || s[0] == '!' || s[0] == '~' || s[0] == ';');
    if (is_operator) {
      if (s[0] == '/' && tk.size() >= 2 && tk[tk.size() - 1].k == NUMB && tk[tk.size() - 2].k == NUMB) {
        tk.pop_back();
        tk.back().small = false;                      // NUM NUM / -> NUM
      } else if (s[0] == '/' && tk.size() >= 2 && tk[tk.size() - 1].k == NUMB && tk[tk.size() - 2].k == VARI) {
        tk.back().small = false;
        tk.push_back(T{OPR, '*', 0, false});          // VAR NUM / -> VAR (1/NUM) *
      } else {
        tk.push_back(T{OPR, s[0], 0, false});
      }
    } else {
      const bool sgn = s.size() <= 18 && !s.empty() && (s[0] == '-' || s[0] == '+');
      const std::string name = sgn ? s.substr(1) : s;
      if (s.size() <= 18 && var_index.count(name)) tk.push_back(T{VARI, 0, 0, false});
      else {
        uint64_t v = 0;
        bool small = s.size() <= 18 && !(sgn && s[0] == '-');
        if (small) v = leading_digits(s, sgn ? 1 : 0);
        tk.push_back(T{NUMB, 0, v, small});
      }
    }
  }
  uint64_t depth = 0;
  for (size_t i = 0; i < tk.size(); ++i) {
    const T& t = tk[i];
    ++work.tokens;
    if (t.k != OPR) { ++depth; work.max_rpn_depth = std::max(work.max_rpn_depth, depth); continue; }
    if (depth) --depth;
    auto chain = [&](uint64_t e) -> uint64_t {
      if (e <= 1) return 0;
      return (uint64_t)(63 - __builtin_clzll(e)) + (uint64_t)__builtin_popcountll(e) - 1;
    };
    const bool lit_exp = i > 0 && tk[i - 1].k == NUMB && tk[i - 1].small;
    switch (t.op) {
      case '+': case '-': ++work.ma; break;
      case '*': ++work.mm; break;
      case '/': ++work.mm; ++work.inv; break;
      case '^': work.mm += lit_exp ? chain(tk[i - 1].lit) : 94; break;          // generic exponent: ~1.5*63
      case '!': work.mm += lit_exp ? chain(tk[i - 1].lit) : 94; ++work.ma; break;
      case '~': work.mm += lit_exp ? chain(tk[i - 1].lit) : 94; ++work.inv; break;
      case ';': work.mm += lit_exp ? chain(tk[i - 1].lit) : 94; ++work.inv; ++work.ma; break;
    }
  }
}

// ------------------------------------------------------------------------------------------
// 7. code generation
// ------------------------------------------------------------------------------------------
// The expression tree is first turned into a DAG by hash-consing (structurally equal sub-trees get
// one representative; x+y and y+x are equal).  Code is then generated depth first for the
// accumulator machine:
//   * a shared compound value is computed at its first use, parked in a TEMP slot (OP_SAVE) and
//     read back (SRC_TMP) at the other uses, as long as temp slots are available;
//   * two sibling sums that share compound terms (f1 of the benchmarks: the numerator and the
//     denominator are sums over the same twenty powers) are evaluated together: every distinct term
//     is computed once and accumulated into one temp slot per sum (OP_ACCUM);
//   * u*u becomes OP_SQR.
// All of this preserves the value of every node exactly (field operations are exact), it only
// avoids recomputing equal residues.
namespace {
constexpr int MAX_TEMPS = 8;
constexpr int SLOT_BUDGET = 6;

struct DagKey {
  uint8_t kind;
  int32_t a, b;
  uint32_t extra;
  bool operator==(const DagKey& o) const { return kind == o.kind && a == o.a && b == o.b && extra == o.extra; }
};
struct DagKeyHash {
  size_t operator()(const DagKey& k) const {
    uint64_t h = 0x9E3779B97F4A7C15ULL * (k.kind + 1);
    h ^= (uint64_t)(uint32_t)k.a * 0xBF58476D1CE4E5B9ULL;
    h = (h << 13) | (h >> 51);
    h ^= (uint64_t)(uint32_t)k.b * 0x94D049BB133111EBULL;
    h = (h << 17) | (h >> 47);
    h ^= k.extra;
    return (size_t)(h * 0x2545F4914F6CDD1DULL);
  }
};

struct FnEmitter {
  const Function& f;
  uint32_t fn;
  std::vector<Instr>& tape;
  std::vector<ConstPatch>& patches;
  bool enable_cse;

  std::vector<int32_t> rep;        // node -> representative
  std::vector<int32_t> remaining;  // representative -> consumptions still expected
  std::vector<int32_t> temp_of;    // representative -> temp slot holding its value, or -1
  std::vector<int32_t> pins;       // representative -> scheduled direct uses that must still find the temp
  std::vector<uint8_t> computed;   // representative -> has been brought into ACC at least once
  bool failed = false;             // an operand was expected in a temp slot that no longer holds it
  bool slot_busy[MAX_TEMPS] = {};
  int max_temps = MAX_TEMPS;       // temp slots this function may use
  uint32_t depth = 0, max_depth = 0, temps_used = 0;
  int32_t root;                    // what this emitter evaluates: the function's root, or only its numerator
  std::vector<Instr> raw;          // instructions scheduled verbatim (EMIT items)

  // den_mode: how the function's top-level division is served (Collection::finish, shared denominators)
  //   DEN_NONE      the whole expression, as written
  //   DEN_PRODUCE   1/D is computed once, parked in temp slot 0 for the functions that follow, then N * that
  //   DEN_CONSUME   N * (temp slot 0): the function before this one has the same denominator
  FnEmitter(const Function& f_, uint32_t fn_, std::vector<Instr>& t, std::vector<ConstPatch>& p, bool cse,
            int temp_budget, int den_mode = 0)
      : f(f_), fn(fn_), tape(t), patches(p), enable_cse(cse) {
    root = den_mode == 2 ? f.nodes[f.root].l : f.root;
    // Spill and temp slots share the kernel's shared memory ([slot][probe][thread], 16 KB per slot at
    // 8 probes x 256 threads): more than SLOT_BUDGET slots per lane cost resident blocks.  Functions that
    // already spill deep get fewer temps; two are always allowed (co-evaluated sums need two).
    if (temp_budget > 0) max_temps = std::min(temp_budget, MAX_TEMPS);
    else max_temps = std::max(2, std::min(MAX_TEMPS, SLOT_BUDGET - (int)f.nodes[root].need));
    if (den_mode) {
      slot_busy[0] = true;         // temp slot 0 carries 1/D from function to function
      temps_used = 1;                // it comes out of the function's own budget: the slot count (occupancy) stays
    }
    build_dag();
  }

  // ---- DAG ----------------------------------------------------------------------------------
  void build_dag() {
    const size_t n = f.nodes.size();
    rep.resize(n);
    remaining.assign(n, 0);
    temp_of.assign(n, -1);
    pins.assign(n, 0);
    computed.assign(n, 0);
    if (!enable_cse) {
      for (size_t i = 0; i < n; ++i) { rep[i] = (int32_t)i; remaining[i] = 1; }
      return;
    }
    std::unordered_map<DagKey, int32_t, DagKeyHash> seen;
    std::unordered_map<std::string, uint32_t> lit_id;
    seen.reserve(n * 2);
    for (size_t i = 0; i < n; ++i) {
      const Node& nd = f.nodes[i];
      DagKey k{(uint8_t)nd.kind, -1, -1, 0};
      if (nd.kind == N_VAR || nd.kind == N_NEGVAR) k.extra = nd.var;
      else if (nd.kind == N_NUM) k.extra = lit_id.emplace(f.literals[nd.lit].text, (uint32_t)lit_id.size()).first->second;
      else {
        k.a = rep[nd.l];
        k.b = rep[nd.r];
        if ((nd.kind == N_ADD || nd.kind == N_MUL) && k.a > k.b) std::swap(k.a, k.b);
      }
      rep[i] = seen.emplace(k, (int32_t)i).first->second;
    }
    // references from distinct reachable parents
    std::vector<char> visited(n, 0);
    std::vector<int32_t> st{rep[root]};
    visited[rep[root]] = 1;
    remaining[rep[root]] = 1;
    while (!st.empty()) {
      const int32_t u = st.back();
      st.pop_back();
      const Node& nd = f.nodes[u];
      if (nd.l < 0) continue;
      for (int32_t c : {rep[nd.l], rep[nd.r]}) {
        ++remaining[c];
        if (!visited[c]) { visited[c] = 1; st.push_back(c); }
      }
    }
  }

  const Node& node(int32_t i) const { return f.nodes[rep[i]]; }
  bool leaf_like(int32_t i) const {
    const Node& n = node(i);
    return n.is_const || n.kind == N_VAR || n.kind == N_NEGVAR;
  }
  bool direct(int32_t i) const { return leaf_like(i) || temp_of[rep[i]] >= 0; }
  bool shared(int32_t i) const { return enable_cse && !leaf_like(i) && remaining[rep[i]] >= 2; }

  int alloc_temp() {
    for (int s = 0; s < max_temps; ++s)
      if (!slot_busy[s]) {
        slot_busy[s] = true;
        temps_used = std::max<uint32_t>(temps_used, (uint32_t)s + 1);
        return s;
      }
    return -1;
  }
  // one expected consumption of node i has happened
  void consume(int32_t i) {
    const int32_t u = rep[i];
    --remaining[u];
    maybe_release(u);
  }
  void maybe_release(int32_t u) {
    if (remaining[u] <= 0 && pins[u] == 0 && temp_of[u] >= 0) {
      slot_busy[temp_of[u]] = false;
      temp_of[u] = -1;
    }
  }
  // a work item scheduled now will read node i straight from its temp slot later: keep the slot until then
  void pin(int32_t i) {
    if (!leaf_like(i)) ++pins[rep[i]];
  }
  void unpin(int32_t i) {
    if (!leaf_like(i)) --pins[rep[i]];
  }

  static Instr mk(uint8_t op, uint8_t flags = 0) {
    Instr i{};
    i.op = op;
    i.flags = flags;
    return i;
  }
  void put(Instr ins) { tape.push_back(ins); }
  // operand fields from a directly addressable node (consumes it)
  void operand(Instr& ins, int32_t i, bool inverse = false) {
    const int32_t u = rep[i];
    const Node& n = f.nodes[u];
    if (n.is_const) {
      ins.src = SRC_CONST;
      patches.push_back(ConstPatch{(uint32_t)tape.size(), fn, u, inverse, false});
    } else if (n.kind == N_VAR || n.kind == N_NEGVAR) {
      ins.src = n.kind == N_VAR ? SRC_VAR : SRC_NEGVAR;
      ins.arg = n.var;
    } else {
      if (temp_of[u] < 0) failed = true;
      ins.src = SRC_TMP;
      ins.arg = (uint32_t)temp_of[u];
      consume(u);
    }
  }

  // ---- sums -----------------------------------------------------------------------------------
  struct Term { int32_t node; bool minus; };
  bool is_sum(int32_t i) const {
    const NodeKind k = node(i).kind;
    return !leaf_like(i) && (k == N_ADD || k == N_SUB);
  }
  // signed terms of the maximal unshared +/- chain below sum root i
  void flatten(int32_t i, std::vector<Term>& out) const {
    std::vector<Term> st{{rep[i], false}};
    bool first = true;
    while (!st.empty()) {
      const Term t = st.back();
      st.pop_back();
      const Node& n = f.nodes[t.node];
      const bool expand = !leaf_like(t.node) && (n.kind == N_ADD || n.kind == N_SUB) && temp_of[t.node] < 0 &&
                          (first || remaining[t.node] == 1);
      first = false;
      if (!expand) { out.push_back(t); continue; }
      // right pushed first so that the left term comes out first
      st.push_back(Term{rep[n.r], n.kind == N_SUB ? !t.minus : t.minus});
      st.push_back(Term{rep[n.l], t.minus});
    }
  }
  bool reaches(int32_t from, int32_t target) const {
    std::vector<int32_t> st{rep[from]};
    std::vector<char> seen(f.nodes.size(), 0);
    seen[rep[from]] = 1;
    while (!st.empty()) {
      const int32_t u = st.back();
      st.pop_back();
      if (u == rep[target]) return true;
      const Node& n = f.nodes[u];
      if (n.l < 0) continue;
      for (int32_t c : {rep[n.l], rep[n.r]})
        if (!seen[c]) { seen[c] = 1; st.push_back(c); }
    }
    return false;
  }

  // ---- generation -------------------------------------------------------------------------------
  struct Item {
    uint8_t what;
    int32_t node;
    bool push, swap;
    int32_t other;
    int32_t aux;
  };
  enum { GEN, AFTER, APPLY_POP, APPLY_DIRECT, INVERT, POW_CONST, MUL_INVCONST, SQUARE, PRE, ACCUM, ADOPT, ADDPOP, EMIT };
  std::vector<Item> work;
  // accumulate lists of co-evaluated sums: for ACCUM items
  struct Acc { int slot; bool minus; bool init; };
  struct AccList { std::vector<Acc> acc; bool post_consume = false; bool emit = true; };
  std::vector<AccList> acc_lists;

  void apply(const Node& n, bool acc_is_right, bool from_stack, int32_t other) {
    auto finish = [&](Instr ins) {
      if (from_stack) { ins.src = SRC_POP; put(ins); --depth; }
      else { unpin(other); operand(ins, other); put(ins); }
    };
    switch (n.kind) {
      case N_ADD: finish(mk(OP_ADD)); break;
      case N_MUL: finish(mk(OP_MUL)); break;
      case N_SUB: finish(mk(OP_SUB, acc_is_right ? F_SWAP : 0)); break;
      case N_DIV: finish(mk(OP_MUL)); break;           // the divisor has been inverted already
      case N_POW: finish(mk(OP_POWL, acc_is_right ? F_SWAP : 0)); break;
      case N_POWNEG: finish(mk(OP_POWL, (acc_is_right ? F_SWAP : 0) | F_NEGR)); break;
      case N_NEGPOW: finish(mk(OP_NPOWL, acc_is_right ? F_SWAP : 0)); break;
      case N_NEGPOWNEG: finish(mk(OP_NPOWL, (acc_is_right ? F_SWAP : 0) | F_NEGR)); break;
      default: break;
    }
  }

  // Two sibling sums sharing compound terms are evaluated in one sweep over their distinct terms.
  struct Plan {
    std::vector<int32_t> order;                               // distinct terms, first appearance first
    std::vector<std::vector<std::pair<int, bool>>> uses;      // per term: (which sum, minus)
  } plan;
  bool can_coevaluate(int32_t l, int32_t r) {
    if (!enable_cse || !is_sum(l) || !is_sum(r) || rep[l] == rep[r]) return false;
    if (temp_of[rep[l]] >= 0 || temp_of[rep[r]] >= 0) return false;
    std::vector<Term> tl, tr;
    flatten(l, tl);
    flatten(r, tr);
    if (tl.size() + tr.size() > 4096) return false;
    plan.order.clear();
    plan.uses.clear();
    std::unordered_map<int32_t, size_t> pos;
    bool any_shared_compound = false;
    auto add = [&](const std::vector<Term>& ts, int which) {
      for (const Term& t : ts) {
        auto it = pos.find(t.node);
        if (it == pos.end()) {
          it = pos.emplace(t.node, plan.order.size()).first;
          plan.order.push_back(t.node);
          plan.uses.emplace_back();
        } else if (which == 1 && !leaf_like(t.node)) {
          for (auto& u : plan.uses[it->second]) any_shared_compound |= (u.first == 0);
        }
        plan.uses[it->second].push_back({which, t.minus});
      }
    };
    add(tl, 0);
    add(tr, 1);
    if (!any_shared_compound) return false;
    if (reaches(l, r) || reaches(r, l)) return false;        // neither sum may live inside the other
    int free_slots = 0;
    for (int s = 0; s < max_temps; ++s) free_slots += !slot_busy[s];
    return free_slots >= 2;
  }
  void schedule_coevaluation(int32_t l, int32_t r, bool push) {
    const int sl = alloc_temp(), sr = alloc_temp();
    // reverse execution order: the two ADOPTs run last
    work.push_back(Item{ADOPT, rep[r], false, false, -1, sr});
    work.push_back(Item{ADOPT, rep[l], false, false, -1, sl});
    // Terms are grouped by HOW the two sums use them -- (+,+), (+,-), only the first sum, ... -- and every
    // group is summed in the accumulator like an ordinary chain before it is added to (or subtracted from)
    // the sums' temp slots: one pair of ACCUMs per group instead of per term.  f1: numerator = A + U,
    // denominator = A - U with A the shared first ten powers and U the sum of the other ten.  Addition in
    // Z_p is associative and commutative, the sums keep their values.
    typedef std::vector<std::pair<int, bool>> Pattern;
    std::vector<Pattern> patterns;
    std::vector<std::vector<size_t>> members;
    for (size_t k = 0; k < plan.order.size(); ++k) {
      Pattern pat = plan.uses[k];
      std::sort(pat.begin(), pat.end());
      bool simple = true;                       // a term used twice by the same sum stays on its own
      for (size_t q = 1; q < pat.size(); ++q) simple &= pat[q].first != pat[q - 1].first;
      size_t g = patterns.size();
      if (simple)
        for (size_t q = 0; q < patterns.size(); ++q)
          if (patterns[q] == pat && !members[q].empty() && plan.uses[members[q][0]].size() == pat.size()) { g = q; break; }
      if (g == patterns.size() || !simple) {
        g = patterns.size();
        patterns.push_back(simple ? pat : Pattern{{-1 - (int)k, false}});   // unique key for the odd ones
        members.emplace_back();
      }
      members[g].push_back(k);
    }
    bool init[2] = {true, true};
    std::vector<Item> fwd;
    bool first_overall = true;
    for (size_t g = 0; g < members.size(); ++g) {
      for (size_t m = 0; m < members[g].size(); ++m) {
        const size_t k = members[g][m];
        AccList uses_only;                       // bookkeeping of this term's uses (nothing emitted)
        uses_only.emit = false;
        uses_only.acc.resize(plan.uses[k].size());
        acc_lists.push_back(std::move(uses_only));
        const int32_t li = (int32_t)acc_lists.size() - 1;
        fwd.push_back(Item{PRE, plan.order[k], false, false, -1, li});
        fwd.push_back(Item{GEN, plan.order[k], first_overall ? push : m > 0, false, -1, 0});
        fwd.push_back(Item{ACCUM, plan.order[k], false, false, -1, li});
        if (m > 0) fwd.push_back(Item{ADDPOP, -1, false, false, -1, 0});
        first_overall = false;
      }
      AccList lst;                               // the group's total goes to the sums
      for (auto& u : plan.uses[members[g][0]]) {
        lst.acc.push_back(Acc{u.first == 0 ? sl : sr, u.second, init[u.first]});
        init[u.first] = false;
      }
      acc_lists.push_back(std::move(lst));
      fwd.push_back(Item{ACCUM, -1, false, false, -1, (int32_t)acc_lists.size() - 1});
    }
    for (size_t k = fwd.size(); k-- > 0;) work.push_back(fwd[k]);
  }

  // ACC = N * (temp slot 0), the parked inverse of the denominator
  void gen_consume() {
    Instr m = mk(OP_MUL);
    m.src = SRC_TMP;
    m.arg = 0;
    raw.push_back(m);
    work.push_back(Item{EMIT, -1, false, false, -1, (int32_t)raw.size() - 1});
    gen(root);
  }
  // temp slot 0 = 1/D (for this function and the ones after it), then ACC = N * that
  void gen_produce() {
    const Node& n = f.nodes[rep[root]];
    const int32_t l = rep[n.l], r = rep[n.r];
    --remaining[rep[root]];
    Instr m = mk(OP_MUL);
    m.src = SRC_TMP;
    m.arg = 0;
    raw.push_back(m);
    Instr sv = mk(OP_SAVE);
    sv.arg = 0;
    raw.push_back(sv);
    // reverse execution order
    work.push_back(Item{EMIT, -1, false, false, -1, (int32_t)raw.size() - 2});
    work.push_back(Item{GEN, l, false, false, -1, 0});
    work.push_back(Item{EMIT, -1, false, false, -1, (int32_t)raw.size() - 1});
    work.push_back(Item{INVERT, -1, false, false, -1, 0});
    gen(r);
  }
  void gen(int32_t root) {
    work.push_back(Item{GEN, root, false, false, -1, 0});
    while (!work.empty()) {
      const Item it = work.back();
      work.pop_back();
      switch (it.what) {
        case INVERT: put(mk(OP_INV)); continue;
        case SQUARE: put(mk(OP_SQR)); continue;
        case EMIT: put(raw[it.aux]); continue;
        case APPLY_POP: apply(node(it.node), it.swap, true, -1); continue;
        case APPLY_DIRECT: apply(node(it.node), it.swap, false, it.other); continue;
        case POW_CONST: {
          const Node& n = node(it.node);
          Instr ins = mk((n.kind == N_POW || n.kind == N_POWNEG) ? OP_POWU : OP_NPOWU,
                         (n.kind == N_POWNEG || n.kind == N_NEGPOWNEG) ? F_NEGR : 0);
          ins.src = SRC_CONST;
          patches.push_back(ConstPatch{(uint32_t)tape.size(), fn, rep[n.r], false, false});
          put(ins);
          continue;
        }
        case MUL_INVCONST: {
          const Node& n = node(it.node);
          Instr ins = mk(OP_MUL);
          ins.src = SRC_CONST;
          patches.push_back(ConstPatch{(uint32_t)tape.size(), fn, rep[n.r], true, false});
          put(ins);
          continue;
        }
        case AFTER: {
          // ACC holds shared node it.node, consumed once by the parent that asked for it
          const int32_t u = rep[it.node];
          --remaining[u];
          if (remaining[u] >= 1 && temp_of[u] < 0) {
            const int s = alloc_temp();
            if (s >= 0) {
              Instr ins = mk(OP_SAVE);
              ins.arg = (uint32_t)s;
              put(ins);
              temp_of[u] = s;
            }
          }
          continue;
        }
        case PRE: {
          // the block serves all its references to this term with one evaluation
          AccList& al = acc_lists[it.aux];
          const int32_t t = rep[it.node];
          const int extra = (int)al.acc.size() - 1;
          if (leaf_like(t) || extra <= 0) continue;
          if (temp_of[t] >= 0) al.post_consume = true;       // read from its temp: release after the load
          else remaining[t] -= extra;
          continue;
        }
        case ADDPOP: {
          Instr ins = mk(OP_ADD);
          ins.src = SRC_POP;
          put(ins);
          --depth;
          continue;
        }
        case ACCUM: {
          if (acc_lists[it.aux].post_consume)
            for (size_t k = 1; k < acc_lists[it.aux].acc.size(); ++k) consume(it.node);
          if (!acc_lists[it.aux].emit) continue;
          for (const Acc& a : acc_lists[it.aux].acc) {
            Instr ins = mk(OP_ACCUM, (a.minus ? F_SWAP : 0) | (a.init ? F_INIT : 0));
            ins.arg = (uint32_t)a.slot;
            put(ins);
          }
          continue;
        }
        case ADOPT: {
          // the sum's value now lives in temp slot it.aux
          temp_of[it.node] = it.aux;
          continue;
        }
        default: break;
      }
      // ---- GEN: bring node it.node into ACC
      const int32_t u = rep[it.node];
      const Node& n = f.nodes[u];
      if (direct(u)) {
        Instr ins = mk(OP_LOAD, it.push ? F_PUSH : 0);
        if (it.push) { ++depth; max_depth = std::max(max_depth, depth); }
        operand(ins, u);
        put(ins);
        continue;
      }
      const int32_t l = rep[n.l], r = rep[n.r];
      const bool is_pow = n.kind == N_POW || n.kind == N_POWNEG || n.kind == N_NEGPOW || n.kind == N_NEGPOWNEG;
      if (!is_pow && can_coevaluate(l, r)) {
        // sibling sums with common terms: run the joint block first, then come back to this node,
        // whose operands are then both parked in temps
        work.push_back(Item{GEN, u, false, false, -1, 0});
        schedule_coevaluation(l, r, it.push);
        continue;
      }
      if (computed[u]) {
        // computed before but could not be parked (no free temp slot): this is a recomputation, the
        // operands are consumed once more than the DAG's edge count says
        ++remaining[l];
        ++remaining[r];
      }
      computed[u] = 1;
      if (shared(u)) work.push_back(Item{AFTER, u, false, false, -1, 0});
      else --remaining[u];
      if (is_pow && f.nodes[r].is_const) {
        work.push_back(Item{POW_CONST, u, false, false, -1, 0});
        work.push_back(Item{GEN, l, it.push, false, -1, 0});
        continue;
      }
      if (n.kind == N_DIV && f.nodes[r].is_const) {
        work.push_back(Item{MUL_INVCONST, u, false, false, -1, 0});
        work.push_back(Item{GEN, l, it.push, false, -1, 0});
        continue;
      }
      if (n.kind == N_MUL && l == r && !direct(l)) {
        work.push_back(Item{SQUARE, u, false, false, -1, 0});
        work.push_back(Item{GEN, l, it.push, false, -1, 0});
        --remaining[l];                               // the second reference is served by the square
        continue;
      }
      const bool push = it.push;
      const bool dl = direct(l), dr = direct(r);
      const unsigned a = f.nodes[l].need, b = f.nodes[r].need;
      const bool left_first = std::max(a, b + 1) <= std::max(b, a + 1);
      if (n.kind == N_DIV) {
        if (dl) {                      // invert the divisor, multiply by the simple numerator
          pin(l);
          work.push_back(Item{APPLY_DIRECT, u, false, true, l, 0});
          work.push_back(Item{INVERT, -1, false, false, -1, 0});
          work.push_back(Item{GEN, r, push, false, -1, 0});
        } else if (left_first) {
          work.push_back(Item{APPLY_POP, u, false, true, -1, 0});
          work.push_back(Item{INVERT, -1, false, false, -1, 0});
          work.push_back(Item{GEN, r, true, false, -1, 0});
          work.push_back(Item{GEN, l, push, false, -1, 0});
        } else {
          work.push_back(Item{APPLY_POP, u, false, false, -1, 0});
          work.push_back(Item{GEN, l, true, false, -1, 0});
          work.push_back(Item{INVERT, -1, false, false, -1, 0});
          work.push_back(Item{GEN, r, push, false, -1, 0});
        }
      } else if (dr) {
        pin(r);
        work.push_back(Item{APPLY_DIRECT, u, false, false, r, 0});
        work.push_back(Item{GEN, l, push, false, -1, 0});
      } else if (dl) {
        pin(l);
        work.push_back(Item{APPLY_DIRECT, u, false, true, l, 0});
        work.push_back(Item{GEN, r, push, false, -1, 0});
      } else if (left_first) {
        work.push_back(Item{APPLY_POP, u, false, true, -1, 0});
        work.push_back(Item{GEN, r, true, false, -1, 0});
        work.push_back(Item{GEN, l, push, false, -1, 0});
      } else {
        work.push_back(Item{APPLY_POP, u, false, false, -1, 0});
        work.push_back(Item{GEN, l, true, false, -1, 0});
        work.push_back(Item{GEN, r, push, false, -1, 0});
      }
    }
  }
};
}  // namespace

void Collection::emit_function(uint32_t fn, const std::vector<uint32_t>& out_slots, int den_mode) {
  const size_t tape0 = tape.size(), patch0 = patches.size();
  for (int attempt = 0; attempt < 2; ++attempt) {
    const bool cse = enable_cse && !no_cse_fn_[fn];
    FnEmitter em(efn(fn), fn, tape, patches, cse, temp_budget, den_mode);
    if (den_mode == 1) em.gen_produce();
    else if (den_mode == 2) em.gen_consume();
    else em.gen(efn(fn).root);
    if (em.failed && cse) {
      // the sharing pass lost track of a temp slot: emit this function as a plain tree instead
      tape.resize(tape0);
      patches.resize(patch0);
      no_cse_fn_[fn] = 1;
      ++cse_fallbacks;
      continue;
    }
    for (uint32_t slot : out_slots) {
      Instr st = FnEmitter::mk(OP_STORE);
      st.arg = slot;
      em.put(st);
    }
    stack_depth = std::max(stack_depth, em.max_depth);
    n_temps = std::max(n_temps, em.temps_used);
    n_slots = std::max(n_slots, em.max_depth + em.temps_used);
    return;
  }
}

// Translation check: run the emitted code of function fn on the host at one point and compare with the
// value of its expression tree (eval_nodes).  The tape must be resolved for prime p.
bool Collection::check_function(uint32_t fn, uint64_t p, const uint64_t* point, uint64_t* temps, bool* temp_set) const {
  std::vector<uint64_t> val;
  eval_nodes(functions[fn], p, point, val);
  const uint64_t want = val[functions[fn].root];
  uint64_t acc = 0;
  std::vector<uint64_t> stack;
  // temps / temp_set (16 entries) belong to the caller: temp slot 0 can carry the inverse of a denominator from
  // one function to the next (shared denominators), exactly as the shared-memory slot does on the device
  bool stored = false;
  for (uint32_t pc = fn_start[pos_of_[fn]]; pc < fn_start[pos_of_[fn] + 1]; ++pc) {
    const Instr& in = tape[pc];
    uint64_t b = 0;
    switch (in.op >= OP_TERM ? (uint8_t)SRC_NONE : in.src) {   // the fused forms decode their own operands
      case SRC_POP:
        if (stack.empty()) return false;
        b = stack.back();
        stack.pop_back();
        break;
      case SRC_VAR: b = point[in.arg] % p; break;
      case SRC_NEGVAR: b = ffhost::neg(point[in.arg] % p, p); break;
      case SRC_CONST: b = in.imm; break;
      case SRC_TMP:
        if (in.arg >= 16 || !temp_set[in.arg]) return false;
        b = temps[in.arg];
        break;
      default: break;
    }
    const bool swap = in.flags & F_SWAP;
    switch (in.op) {
      case OP_LOAD:
        if (in.flags & F_PUSH) stack.push_back(acc);
        acc = b;
        break;
      case OP_ADD: acc = ffhost::add(acc, b, p); break;
      case OP_SUB: acc = swap ? ffhost::sub(b, acc, p) : ffhost::sub(acc, b, p); break;
      case OP_MUL: acc = ffhost::mul(acc, b, p); break;
      case OP_SQR: acc = ffhost::mul(acc, acc, p); break;
      case OP_NEG: acc = ffhost::neg(acc, p); break;
      case OP_INV: acc = ffhost::inv(acc, p); break;
      case OP_POWU: acc = ffhost::pow_ff(acc, in.imm, p); break;
      case OP_NPOWU: acc = ffhost::pow_int(acc, ffhost::to_neg_int(in.imm), p); break;
      case OP_POWL: acc = swap ? ffhost::pow_ff(b, acc, p) : ffhost::pow_ff(acc, b, p); break;
      case OP_NPOWL:
        acc = swap ? ffhost::pow_int(b, ffhost::to_neg_int(acc), p) : ffhost::pow_int(acc, ffhost::to_neg_int(b), p);
        break;
      case OP_MULADD: {
        const uint64_t m = ffhost::mul(acc, point[in.arg] % p, p);
        acc = swap ? ffhost::sub(in.imm, m, p) : ffhost::add(m, in.imm, p);
        break;
      }
      case OP_SAVE:
        if (in.arg >= 16) return false;
        temps[in.arg] = acc;
        temp_set[in.arg] = true;
        break;
      case OP_ACCUM: {
        if (in.arg >= 16) return false;
        const uint64_t v = swap ? ffhost::neg(acc, p) : acc;
        if (in.flags & F_INIT) temps[in.arg] = v;
        else if (!temp_set[in.arg]) return false;
        else temps[in.arg] = ffhost::add(temps[in.arg], v, p);
        temp_set[in.arg] = true;
        break;
      }
      case OP_STORE:
        if (acc != want) return false;
        stored = true;
        break;
      case OP_TERM: {
        const uint8_t mode = (in.flags >> TM_SHIFT) & 7;
        uint64_t t = ffhost::pow_ff(point[in.arg & 0xffff] % p, in.arg >> 16, p);
        if (in.flags & F_CMUL) t = ffhost::mul(t, in.imm, p);
        if (mode == TM_LOAD) {
          if (in.flags & F_PUSH) stack.push_back(acc);
          acc = t;
        } else if (mode == TM_ADD) acc = ffhost::add(acc, t, p);
        else if (mode == TM_SUB) acc = ffhost::sub(acc, t, p);
        else if (mode == TM_RSUB) acc = ffhost::sub(t, acc, p);
        else if (mode == TM_MUL) acc = ffhost::mul(acc, t, p);
        else return false;
        continue;                                  // the F_* bits above are not the generic flags
      }
      case OP_ACCUM2: {
        const uint32_t sa = in.arg & 0xff, sb = (in.arg >> 8) & 0xff;
        if (sa >= 16 || sb >= 16) return false;
        const uint64_t va = (in.flags & F_SWAP) ? ffhost::neg(acc, p) : acc;
        const uint64_t vb = (in.flags & F_SWAP2) ? ffhost::neg(acc, p) : acc;
        if (in.flags & F_INIT) temps[sa] = va;
        else if (!temp_set[sa]) return false;
        else temps[sa] = ffhost::add(temps[sa], va, p);
        temp_set[sa] = true;
        if (in.flags & F_INIT2) temps[sb] = vb;
        else if (!temp_set[sb]) return false;
        else temps[sb] = ffhost::add(temps[sb], vb, p);
        temp_set[sb] = true;
        continue;
      }
      case OP_ADDVARS:
      case OP_MULVARS: {
        const uint32_t cnt = in.arg & 0xff, signs = in.arg >> 8;
        for (uint32_t k = 0; k < cnt; ++k) {
          const uint64_t v = point[(in.imm >> (8 * k)) & 0xff] % p;
          if (in.op == OP_MULVARS) acc = ffhost::mul(acc, v, p);
          else acc = ((signs >> k) & 1) ? ffhost::sub(acc, v, p) : ffhost::add(acc, v, p);
        }
        continue;
      }
      default: return false;
    }
    if (in.flags & F_NEGR) acc = ffhost::neg(acc, p);
  }
  return stored && stack.empty();
}

// ------------------------------------------------------------------------------------------
// 8. collection level
// ------------------------------------------------------------------------------------------
namespace {
const uint64_t P0 = 9223372036854775783ULL;   // 2^63 - 25, the largest 63-bit prime
const uint64_t P1 = 9223372036854775643ULL;   // 2^63 - 165, the next one
inline uint64_t splitmix64(uint64_t& s) {
  uint64_t z = (s += 0x9E3779B97F4A7C15ULL);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
  return z ^ (z >> 31);
}
}  // namespace

// Structure of a function for duplicate detection ACROSS the functions of a collection: two independent 64-bit
// fingerprints per node, bottom up (the node array is in post order); x+y and y+x, x*y and y*x are the same.
// Equal fingerprints are confirmed by an exact comparison before anything is shared.
namespace {
struct Fp { uint64_t a, b; };
inline uint64_t mix64(uint64_t x) {
  x ^= x >> 30; x *= 0xBF58476D1CE4E5B9ULL;
  x ^= x >> 27; x *= 0x94D049BB133111EBULL;
  return x ^ (x >> 31);
}
void fingerprints(const Function& f, std::vector<Fp>& fp) {
  fp.resize(f.nodes.size());
  for (size_t i = 0; i < f.nodes.size(); ++i) {
    const Node& n = f.nodes[i];
    Fp h;
    if (n.kind == N_VAR || n.kind == N_NEGVAR) {
      h.a = mix64(0x1000 + 2 * (uint64_t)n.var + (n.kind == N_NEGVAR));
      h.b = mix64(h.a ^ 0x9E3779B97F4A7C15ULL);
    } else if (n.kind == N_NUM) {
      uint64_t x = 0xCBF29CE484222325ULL, y = 0x84222325CBF29CE4ULL;
      for (unsigned char ch : f.literals[n.lit].text) {
        x = (x ^ ch) * 0x100000001B3ULL;
        y = mix64(y + ch);
      }
      h.a = mix64(x ^ 0x2000);
      h.b = mix64(y ^ 0x3000);
    } else {
      Fp l = fp[n.l], r = fp[n.r];
      if ((n.kind == N_ADD || n.kind == N_MUL) && (l.a > r.a || (l.a == r.a && l.b > r.b))) std::swap(l, r);
      h.a = mix64(l.a * 0x9E3779B97F4A7C15ULL + mix64(r.a + 0x51 * (uint64_t)n.kind) + (uint64_t)n.kind);
      h.b = mix64(r.b * 0xD6E8FEB86659FD93ULL + mix64(l.b ^ (0x77 * (uint64_t)n.kind)) + 3 * (uint64_t)n.kind);
    }
    fp[i] = h;
  }
}
bool same_structure(const Function& f, const std::vector<Fp>& ff, int32_t fr, const Function& g, const std::vector<Fp>& gf,
                    int32_t gr) {
  std::vector<std::pair<int32_t, int32_t>> st{{fr, gr}};
  while (!st.empty()) {
    const auto [i, j] = st.back();
    st.pop_back();
    const Node& a = f.nodes[i];
    const Node& b = g.nodes[j];
    if (a.kind != b.kind || ff[i].a != gf[j].a || ff[i].b != gf[j].b) return false;
    if (a.kind == N_VAR || a.kind == N_NEGVAR) {
      if (a.var != b.var) return false;
    } else if (a.kind == N_NUM) {
      if (f.literals[a.lit].text != g.literals[b.lit].text) return false;
    } else {
      int32_t al = a.l, ar = a.r, bl = b.l, br = b.r;
      if (a.kind == N_ADD || a.kind == N_MUL) {
        // pair the children by fingerprint
        auto less = [](const Fp& x, const Fp& y) { return x.a < y.a || (x.a == y.a && x.b < y.b); };
        if (less(ff[ar], ff[al])) std::swap(al, ar);
        if (less(gf[br], gf[bl])) std::swap(bl, br);
      }
      st.push_back({al, bl});
      st.push_back({ar, br});
    }
  }
  return true;
}
}  // namespace

// ---- shared rows ---------------------------------------------------------------------------------------------
// Sub-expressions that occur in SEVERAL places of a collection -- the powers x^e of a reduction table, common
// factors of denominators, numerators shared by different functions -- are evaluated once per probe instead of once
// per occurrence: each becomes an extra variable row (variable index n_vars + r), filled by a small collection of its
// own (`pre`, one function per row, over the same variables) that is evaluated before this tape at the same points.
// The functions' tape then reads the row like any other coordinate of the probe.  Unlike temp slots, rows live in
// device memory: they survive the tape's chunk boundaries (grid.y) and cost no shared memory, hence no occupancy.
// (The reference shares nothing between functions except whole duplicates under check_is_equal,
// ShuntingYardParser.cpp:125-140.)  Values are unchanged: equal sub-trees give equal residues in exact arithmetic, and
// the translation check in finish() compares every function's tape -- rows included -- with its ORIGINAL expression.
namespace {
uint32_t literal_exponent(const Function& f, const Node& n) {      // 0: not a small plain literal
  if (n.kind != N_NUM) return 0;
  const std::string& t = f.literals[n.lit].text;
  if (t.empty() || t.size() > 6) return 0;
  for (char ch : t)
    if (ch < '0' || ch > '9') return 0;
  return (uint32_t)std::min<uint64_t>(leading_digits(t, 0), 0xffffffu);
}
// multiplications one evaluation of node i costs by itself (its operands not counted); inversions count as 6
uint32_t own_cost(const Function& f, const Node& n) {
  auto chain = [](uint64_t e) -> uint32_t {
    return e <= 1 ? 0u : (uint32_t)(63 - __builtin_clzll(e)) + (uint32_t)__builtin_popcountll(e) - 1u;
  };
  if (n.is_const) return 0;
  switch (n.kind) {
    case N_MUL: return 1;
    case N_DIV: return f.nodes[n.r].is_const ? 1 : 7;
    case N_POW: case N_POWNEG: {
      const uint32_t e = literal_exponent(f, f.nodes[n.r]);
      return e ? chain(e) : 94;
    }
    case N_NEGPOW: case N_NEGPOWNEG: {
      const uint32_t e = literal_exponent(f, f.nodes[n.r]);
      return e ? chain(e) + 6 : 100;
    }
    default: return 0;
  }
}
struct FpKey {
  uint64_t a, b;
  bool operator==(const FpKey& o) const { return a == o.a && b == o.b; }
};
struct FpKeyHash {
  size_t operator()(const FpKey& k) const { return (size_t)(k.a ^ mix64(k.b)); }
};
// copy the part of f reachable from `root` into g; nodes for which `leaf_var(i)` returns a variable index are
// replaced by that variable (and not descended into).  Children precede parents in both arrays.
template <class LeafVar>
void copy_tree(const Function& f, int32_t root, Function& g, LeafVar leaf_var) {
  std::vector<int32_t> reach, st{root};
  std::vector<int32_t> as_var(f.nodes.size(), -2);       // -2 unseen, -1 kept, >= 0 variable
  while (!st.empty()) {
    const int32_t i = st.back();
    st.pop_back();
    if (as_var[i] != -2) continue;                       // (a tree: cannot happen; a DAG-shaped input would)
    const int64_t v = leaf_var(i);
    as_var[i] = v >= 0 ? (int32_t)v : -1;
    reach.push_back(i);
    if (v >= 0 || f.nodes[i].l < 0) continue;
    st.push_back(f.nodes[i].l);
    st.push_back(f.nodes[i].r);
  }
  std::sort(reach.begin(), reach.end());
  std::vector<int32_t> remap(f.nodes.size(), -1);
  g.nodes.clear();
  g.literals.clear();
  g.nodes.reserve(reach.size());
  for (int32_t i : reach) {
    Node n = f.nodes[i];
    if (as_var[i] >= 0) {
      n = Node{};
      n.kind = N_VAR;
      n.is_const = false;
      n.var = (uint32_t)as_var[i];
      n.l = n.r = -1;
    } else if (n.kind == N_NUM) {
      n.lit = (uint32_t)g.literals.size();
      g.literals.push_back(f.literals[f.nodes[i].lit]);
    } else if (n.l >= 0) {
      n.l = remap[n.l];
      n.r = remap[n.r];
    }
    remap[i] = (int32_t)g.nodes.size();
    g.nodes.push_back(n);
  }
  g.root = remap[root];
}
}  // namespace

// ---- long sums ---------------------------------------------------------------------------------------------------
// A function that is one long sum (an amplitude, a reduction coefficient with 10^4 terms) is a serial walk for ONE
// thread block per tile of points: at the Reconstructor's bunch sizes (<= 128 points) the rest of the chip idles.  Its
// sums are cut into partial sums of about SPLIT_TARGET tree nodes each; every partial sum becomes a shared row of
// level 1 (evaluated by the rows' tape, whose chunks run on different SMs), the function adds the rows up.  Addition
// in Z_p is associative and commutative: the value is unchanged.
namespace {
constexpr uint32_t GROUP_BASE = 1u << 24;      // placeholder variable index of partial sum g until the row count is known
uint32_t env_u32(const char* name, uint32_t dflt) {
  const char* e = std::getenv(name);
  return e && *e ? (uint32_t)std::max(1L, std::atol(e)) : dflt;
}
const uint32_t SPLIT_MIN = env_u32("FFGPU_SPLIT_MIN", 768);          // tree nodes below which a sum is left alone
const uint32_t SPLIT_TARGET = env_u32("FFGPU_SPLIT_TARGET", 320);    // tree nodes per partial sum
const uint32_t SPLIT_MAX_GROUPS = env_u32("FFGPU_SPLIT_GROUPS", 96); // partial sums per sum at most

struct Mono {
  std::vector<int32_t> consts;                          // constant sub-trees of the source that multiply the term
  std::vector<std::pair<uint32_t, uint32_t>> pw;        // (variable, exponent), exponent >= 1
  int32_t opaque = -1;                                  // not a monomial: the term's sub-tree, taken as it is
  bool minus = false;
};
// term -> monomial; false if the term is something else
bool parse_monomial(const Function& f, int32_t t, Mono& m) {
  std::vector<int32_t> st{t};
  while (!st.empty()) {
    const int32_t i = st.back();
    st.pop_back();
    const Node& n = f.nodes[i];
    if (n.is_const) { m.consts.push_back(i); continue; }
    if (n.kind == N_VAR) { m.pw.push_back({n.var, 1}); continue; }
    if (n.kind == N_NEGVAR) { m.pw.push_back({n.var, 1}); m.minus = !m.minus; continue; }
    if (n.kind == N_MUL) { st.push_back(n.r); st.push_back(n.l); continue; }
    if ((n.kind == N_POW || n.kind == N_POWNEG) && f.nodes[n.l].kind == N_VAR) {
      const uint32_t e = literal_exponent(f, f.nodes[n.r]);
      if (e >= 1 && e < 65536) {
        m.pw.push_back({f.nodes[n.l].var, e});
        if (n.kind == N_POWNEG) m.minus = !m.minus;
        continue;
      }
    }
    return false;
  }
  std::sort(m.pw.begin(), m.pw.end());
  size_t w = 0;
  for (size_t k = 0; k < m.pw.size(); ++k) {
    if (w && m.pw[w - 1].first == m.pw[k].first) m.pw[w - 1].second += m.pw[k].second;
    else m.pw[w++] = m.pw[k];
  }
  m.pw.resize(w);
  for (auto& q : m.pw)
    if (q.second >= 65536) return false;
  return true;
}
// append the sub-tree of f below `root` to g (children first); returns the index of its root in g
int32_t append_subtree(const Function& f, int32_t root, Function& g, std::vector<int32_t>& remap) {
  std::vector<int32_t> reach, st{root};
  while (!st.empty()) {
    const int32_t i = st.back();
    st.pop_back();
    reach.push_back(i);
    if (f.nodes[i].l < 0) continue;
    st.push_back(f.nodes[i].l);
    st.push_back(f.nodes[i].r);
  }
  std::sort(reach.begin(), reach.end());
  for (int32_t i : reach) {
    Node n = f.nodes[i];
    if (n.kind == N_NUM) {
      n.lit = (uint32_t)g.literals.size();
      g.literals.push_back(f.literals[f.nodes[i].lit]);
    } else if (n.l >= 0) {
      n.l = remap[n.l];
      n.r = remap[n.r];
    }
    remap[i] = (int32_t)g.nodes.size();
    g.nodes.push_back(n);
  }
  return remap[root];
}
int32_t push_binary(Function& g, NodeKind kind, int32_t l, int32_t r) {
  Node n{};
  n.kind = kind;
  n.l = l;
  n.r = r;
  n.is_const = g.nodes[l].is_const && g.nodes[r].is_const;
  g.nodes.push_back(n);
  return (int32_t)g.nodes.size() - 1;
}
int32_t push_var(Function& g, uint32_t var) {
  Node n{};
  n.kind = N_VAR;
  n.var = var;
  n.l = n.r = -1;
  n.is_const = false;
  g.nodes.push_back(n);
  return (int32_t)g.nodes.size() - 1;
}
// Cut the long sums of f into partial sums (appended to `groups`); `out` becomes f with each of them replaced by the
// sum of its placeholder variables.  Returns false (and leaves `out` empty) if f has no long sum.
bool split_long_sums(const Function& f, std::vector<Function>& groups, Function& out) {
  const size_t n = f.nodes.size();
  if (n < SPLIT_MIN) return false;
  std::vector<uint32_t> size(n, 1);
  for (size_t i = 0; i < n; ++i)
    if (f.nodes[i].l >= 0) size[i] = 1 + size[f.nodes[i].l] + size[f.nodes[i].r];
  auto is_sum = [&](int32_t i) { return !f.nodes[i].is_const && (f.nodes[i].kind == N_ADD || f.nodes[i].kind == N_SUB); };
  // the long sums: the outermost ones, wherever they sit (numerator, denominator, a factor, the base of a power)
  std::vector<int32_t> tops;
  {
    std::vector<int32_t> st{f.root};
    while (!st.empty()) {
      const int32_t i = st.back();
      st.pop_back();
      if (size[i] < SPLIT_MIN || f.nodes[i].l < 0 || f.nodes[i].is_const) continue;
      if (is_sum(i)) { tops.push_back(i); continue; }
      st.push_back(f.nodes[i].l);
      st.push_back(f.nodes[i].r);
    }
    std::sort(tops.begin(), tops.end());
  }
  struct Term { int32_t node; bool minus; };
  std::unordered_map<int32_t, std::pair<uint32_t, uint32_t>> repl;       // sum node -> (first group, count)
  std::vector<int32_t> remap(n, -1);
  for (int32_t top : tops) {
    std::vector<Term> terms, st{{top, false}};
    while (!st.empty()) {
      const Term t = st.back();
      st.pop_back();
      if (!is_sum(t.node)) { terms.push_back(t); continue; }
      const Node& nd = f.nodes[t.node];
      st.push_back(Term{nd.r, nd.kind == N_SUB ? !t.minus : t.minus});      // right first: the left term comes out first
      st.push_back(Term{nd.l, t.minus});
    }
    // monomials that look alike (same variables, neighbouring exponents) go into the same partial sum: that is where
    // the Horner nests and the shared rows find something to share.  Other terms keep their order, behind them.
    {
      std::vector<std::pair<std::vector<std::pair<uint32_t, uint32_t>>, Term>> keyed;
      keyed.reserve(terms.size());
      for (const Term& t : terms) {
        Mono m;
        if (!parse_monomial(f, t.node, m)) m.pw.assign(1, {~0u, ~0u});
        keyed.push_back({std::move(m.pw), t});
      }
      std::stable_sort(keyed.begin(), keyed.end(), [](const auto& x, const auto& y) { return x.first < y.first; });
      for (size_t q = 0; q < terms.size(); ++q) terms[q] = keyed[q].second;
    }
    uint64_t total = 0;
    for (const Term& t : terms) total += size[t.node];
    const uint32_t want = (uint32_t)std::min<uint64_t>(SPLIT_MAX_GROUPS, std::max<uint64_t>(2, total / SPLIT_TARGET));
    if (terms.size() < 2) continue;
    const uint64_t per = (total + want - 1) / want;
    const uint32_t first = (uint32_t)groups.size();
    size_t k = 0;
    while (k < terms.size()) {
      Function g;
      uint64_t acc_size = 0;
      int32_t acc = -1;
      std::vector<Term> mine;
      while (k < terms.size() && (mine.empty() || acc_size + size[terms[k].node] / 2 < per)) {
        acc_size += size[terms[k].node];
        mine.push_back(terms[k++]);
      }
      // a term with a plus sign leads; if there is none the sum starts from the literal 0
      size_t lead = 0;
      while (lead < mine.size() && mine[lead].minus) ++lead;
      if (lead == mine.size()) {
        Node z{};
        z.kind = N_NUM;
        z.is_const = true;
        z.l = z.r = -1;
        z.lit = (uint32_t)g.literals.size();
        g.literals.push_back(Literal{"0"});
        g.nodes.push_back(z);
        acc = (int32_t)g.nodes.size() - 1;
      } else {
        acc = append_subtree(f, mine[lead].node, g, remap);
      }
      for (size_t q = 0; q < mine.size(); ++q) {
        if (q == lead) continue;
        const int32_t t = append_subtree(f, mine[q].node, g, remap);
        acc = push_binary(g, mine[q].minus ? N_SUB : N_ADD, acc, t);
      }
      g.root = acc;
      groups.push_back(std::move(g));
    }
    repl[top] = {first, (uint32_t)groups.size() - first};
  }
  if (repl.empty()) return false;
  // f with the long sums replaced by sums of placeholder variables
  std::vector<int32_t> reach, st{f.root};
  while (!st.empty()) {
    const int32_t i = st.back();
    st.pop_back();
    reach.push_back(i);
    if (repl.count(i) || f.nodes[i].l < 0) continue;
    st.push_back(f.nodes[i].l);
    st.push_back(f.nodes[i].r);
  }
  std::sort(reach.begin(), reach.end());
  out = Function{};
  for (int32_t i : reach) {
    auto it = repl.find(i);
    if (it != repl.end()) {
      int32_t acc = push_var(out, GROUP_BASE + it->second.first);
      for (uint32_t q = 1; q < it->second.second; ++q) acc = push_binary(out, N_ADD, acc, push_var(out, GROUP_BASE + it->second.first + q));
      remap[i] = acc;
      continue;
    }
    Node nd = f.nodes[i];
    if (nd.kind == N_NUM) {
      nd.lit = (uint32_t)out.literals.size();
      out.literals.push_back(f.literals[f.nodes[i].lit]);
    } else if (nd.l >= 0) {
      nd.l = remap[nd.l];
      nd.r = remap[nd.r];
    }
    remap[i] = (int32_t)out.nodes.size();
    out.nodes.push_back(nd);
  }
  out.root = remap[f.root];
  return true;
}
}  // namespace

// ---- expanded polynomials -> Horner nests -----------------------------------------------------------------------
// A sum of monomials c * x^a * y^b * ... (what Expand[] leaves behind) costs one multiplication per variable and per
// monomial; factored greedily -- the variable that occurs in most terms first, P = v^m * Q + R, then Q and R the same way
// -- it costs about one per monomial (the FireFly paper measures 5-7x between its Tab. 4 and Tab. 5; the reference has a
// Horner generator for its own output, PolynomialFF.cpp:46-56, but evaluates input as written).  Ring identities only:
// every value is unchanged, constants keep their own nodes and are folded per prime as before.
namespace {
struct HornerBuilder {
  const Function& f;
  Function& g;
  std::vector<int32_t>& remap;
  typedef std::pair<int32_t, bool> Val;                 // (node of g, the value is MINUS that node)
  int32_t literal(const std::string& text) {
    Node z{};
    z.kind = N_NUM;
    z.is_const = true;
    z.l = z.r = -1;
    z.lit = (uint32_t)g.literals.size();
    g.literals.push_back(Literal{text});
    g.nodes.push_back(z);
    return (int32_t)g.nodes.size() - 1;
  }
  int32_t power(uint32_t var, uint32_t e) {
    const int32_t v = push_var(g, var);
    return e == 1 ? v : push_binary(g, N_POW, v, literal(std::to_string(e)));
  }
  Val add(Val a, Val b) {
    if (a.first < 0) return b;
    if (b.first < 0) return a;
    if (!a.second) return {push_binary(g, b.second ? N_SUB : N_ADD, a.first, b.first), false};
    if (!b.second) return {push_binary(g, N_SUB, b.first, a.first), false};
    return {push_binary(g, N_ADD, a.first, b.first), true};
  }
  Val monomial(const Mono& m) {
    if (m.opaque >= 0) return {append_subtree(f, m.opaque, g, remap), m.minus};
    int32_t acc = -1;
    for (int32_t c : m.consts) {
      const int32_t k = append_subtree(f, c, g, remap);
      acc = acc < 0 ? k : push_binary(g, N_MUL, acc, k);
    }
    for (const auto& q : m.pw) {
      const int32_t k = power(q.first, q.second);
      acc = acc < 0 ? k : push_binary(g, N_MUL, acc, k);
    }
    if (acc < 0) acc = literal("1");
    return {acc, m.minus};
  }
  // the signed sum of the monomials, factored
  Val build(std::vector<Mono> terms) {
    Val total{-1, false};
    std::unordered_map<uint32_t, uint32_t> count;
    while (true) {
      if (terms.size() < 2) break;
      count.clear();
      for (const Mono& m : terms)
        if (m.opaque < 0)
          for (const auto& q : m.pw) ++count[q.first];
      uint32_t best = 0, best_n = 1;
      for (const auto& kv : count)
        if (kv.second > best_n || (kv.second == best_n && best_n > 1 && kv.first < best)) { best = kv.first; best_n = kv.second; }
      if (best_n < 2) break;
      uint32_t low = ~0u;
      for (const Mono& m : terms)
        if (m.opaque < 0)
          for (const auto& q : m.pw)
            if (q.first == best) low = std::min(low, q.second);
      std::vector<Mono> inner, rest;
      for (Mono& m : terms) {
        bool has = false;
        if (m.opaque < 0)
          for (size_t k = 0; k < m.pw.size(); ++k)
            if (m.pw[k].first == best) {
              has = true;
              m.pw[k].second -= low;
              if (m.pw[k].second == 0) m.pw.erase(m.pw.begin() + (long)k);
              break;
            }
        (has ? inner : rest).push_back(std::move(m));
      }
      const Val q = build(std::move(inner));                       // depth: the total degree at most
      total = add(total, Val{push_binary(g, N_MUL, power(best, low), q.first), q.second});
      terms = std::move(rest);
    }
    for (const Mono& m : terms) total = add(total, monomial(m));
    return total;
  }
};
// f with every sum of at least three monomials (wherever it sits) replaced by its Horner nest; false: nothing to do
bool hornerize(const Function& f, Function& out) {
  const size_t n = f.nodes.size();
  auto is_sum = [&](int32_t i) { return !f.nodes[i].is_const && (f.nodes[i].kind == N_ADD || f.nodes[i].kind == N_SUB); };
  std::unordered_map<int32_t, std::vector<Mono>> tops;
  {
    std::vector<int32_t> st{f.root};
    while (!st.empty()) {
      const int32_t i = st.back();
      st.pop_back();
      if (f.nodes[i].l < 0 || f.nodes[i].is_const) continue;
      if (is_sum(i)) {
        struct Term { int32_t node; bool minus; };
        std::vector<Term> ts{{i, false}};
        std::vector<Mono> monos;
        size_t n_mono = 0;
        while (!ts.empty()) {
          const Term t = ts.back();
          ts.pop_back();
          if (is_sum(t.node)) {
            const Node& nd = f.nodes[t.node];
            ts.push_back(Term{nd.r, nd.kind == N_SUB ? !t.minus : t.minus});
            ts.push_back(Term{nd.l, t.minus});
            continue;
          }
          Mono m;
          if (!parse_monomial(f, t.node, m)) {
            m = Mono{};
            m.opaque = t.node;
          } else {
            ++n_mono;
          }
          if (t.minus) m.minus = !m.minus;
          monos.push_back(std::move(m));
        }
        if (n_mono >= 3) { tops.emplace(i, std::move(monos)); continue; }
        // too few monomials: look inside the terms instead
        for (const Mono& m : monos)
          if (m.opaque >= 0) st.push_back(m.opaque);
        continue;
      }
      st.push_back(f.nodes[i].l);
      st.push_back(f.nodes[i].r);
    }
  }
  if (tops.empty()) return false;
  std::vector<int32_t> reach, st{f.root}, remap(n, -1);
  while (!st.empty()) {
    const int32_t i = st.back();
    st.pop_back();
    reach.push_back(i);
    if (tops.count(i) || f.nodes[i].l < 0) continue;
    st.push_back(f.nodes[i].l);
    st.push_back(f.nodes[i].r);
  }
  std::sort(reach.begin(), reach.end());
  out = Function{};
  bool changed = false;
  HornerBuilder hb{f, out, remap};
  for (int32_t i : reach) {
    auto it = tops.find(i);
    if (it != tops.end()) {
      // keep the nest only if it multiplies less than the sum as written
      const size_t n0 = out.nodes.size(), l0 = out.literals.size();
      HornerBuilder::Val v = hb.build(std::move(it->second));
      if (v.second) v = {push_binary(out, N_SUB, hb.literal("0"), v.first), false};
      uint64_t after = 0, before = 0;
      for (size_t k = n0; k < out.nodes.size(); ++k) after += own_cost(out, out.nodes[k]);
      std::vector<int32_t> walk{i};
      while (!walk.empty()) {
        const int32_t k = walk.back();
        walk.pop_back();
        before += own_cost(f, f.nodes[k]);
        if (f.nodes[k].l >= 0 && !f.nodes[k].is_const) { walk.push_back(f.nodes[k].l); walk.push_back(f.nodes[k].r); }
      }
      if (after * 10 <= before * 9) {
        remap[i] = v.first;
        changed = true;
      } else {
        out.nodes.resize(n0);
        out.literals.resize(l0);
        remap[i] = append_subtree(f, i, out, remap);
      }
      continue;
    }
    Node nd = f.nodes[i];
    if (nd.kind == N_NUM) {
      nd.lit = (uint32_t)out.literals.size();
      out.literals.push_back(f.literals[f.nodes[i].lit]);
    } else if (nd.l >= 0) {
      nd.l = remap[nd.l];
      nd.r = remap[nd.r];
    }
    remap[i] = (int32_t)out.nodes.size();
    out.nodes.push_back(nd);
  }
  out.root = remap[f.root];
  return changed;
}
}  // namespace

void Collection::build_table(const std::vector<int32_t>& dup_of) {
  lowered_.clear();
  pre.clear();
  n_table_rows = 0;
  const size_t nf = functions.size();
  if (!enable_cse || !enable_cross || !enable_table || nf == 0 || vars.size() >= 60000) return;
  const uint32_t n_base = (uint32_t)vars.size();
  // ---- long sums -> partial sums (level-1 rows)
  std::vector<Function> groups;
  std::vector<Function> cut(nf);               // functions whose long sums were replaced (no nodes: unchanged)
  if (enable_split)
    for (size_t i = 0; i < nf; ++i)
      if (dup_of[i] < 0 && groups.size() < 4 * SPLIT_MAX_GROUPS) split_long_sums(functions[i], groups, cut[i]);
  // expanded polynomials -> Horner nests: in the partial sums, and in the sums of the functions that were not cut
  bool any_cut = !groups.empty();
  if (enable_horner && !small_prime_mode_) {
    Function h;
    for (Function& g : groups)
      if (hornerize(g, h)) g = std::move(h);
    for (size_t i = 0; i < nf; ++i) {
      if (dup_of[i] >= 0) continue;
      if (hornerize(cut[i].nodes.empty() ? functions[i] : cut[i], h)) {
        cut[i] = std::move(h);
        any_cut = true;
      }
    }
  }
  if (nf < 2 && !any_cut) return;
  // the units the sharing pass looks at: every function that is emitted, and the partial sums
  std::vector<const Function*> unit;
  std::vector<int64_t> unit_fn;                // function index, or -1 - g for partial sum g
  for (size_t i = 0; i < nf; ++i)
    if (dup_of[i] < 0) { unit.push_back(cut[i].nodes.empty() ? &functions[i] : &cut[i]); unit_fn.push_back((int64_t)i); }
  for (size_t g = 0; g < groups.size(); ++g) { unit.push_back(&groups[g]); unit_fn.push_back(-1 - (int64_t)g); }
  const size_t nu = unit.size();
  struct Cand {
    uint32_t fn;         // unit of the first occurrence
    int32_t node;
    uint32_t cost;       // multiplications of one evaluation of the whole sub-tree
    uint32_t refs;       // occurrences not covered by a larger selected sub-tree
    bool selected;
    int32_t row;
  };
  std::vector<std::vector<Fp>> fps(nu);
  // Candidates by fingerprint: an open-addressing table of indices into one array (reserved for the most there can
  // be, so its elements never move).  The candidate of every node is looked up once: the passes below walk the trees
  // again and again, and on a 10 000-function collection the lookups were half of the whole compilation.
  std::vector<Cand> cands;
  std::vector<FpKey> cand_key;
  std::vector<uint32_t> cand_slot;             // 0 = empty, else index + 1
  size_t cand_mask = 0;
  {
    size_t inner_nodes = 0;
    for (size_t i = 0; i < nu; ++i)
      for (const Node& n : unit[i]->nodes) inner_nodes += n.l >= 0 && !n.is_const;
    cands.reserve(inner_nodes);
    cand_key.reserve(inner_nodes);
    size_t cap = 16;
    while (cap < 2 * inner_nodes) cap <<= 1;
    cand_slot.assign(cap, 0);
    cand_mask = cap - 1;
  }
  auto cand_emplace = [&](const FpKey& key, const Cand& fresh) -> std::pair<Cand*, bool> {
    size_t h = FpKeyHash{}(key) & cand_mask;
    for (;;) {
      const uint32_t s = cand_slot[h];
      if (!s) {
        cands.push_back(fresh);
        cand_key.push_back(key);
        cand_slot[h] = (uint32_t)cands.size();
        return {&cands.back(), true};
      }
      if (cand_key[s - 1] == key) return {&cands[s - 1], false};
      h = (h + 1) & cand_mask;
    }
  };
  std::vector<std::vector<Cand*>> cand_of(nu);
  std::vector<uint32_t> cost;
  for (size_t i = 0; i < nu; ++i) {
    const Function& f = *unit[i];
    fingerprints(f, fps[i]);
    cost.assign(f.nodes.size(), 0);
    cand_of[i].assign(f.nodes.size(), nullptr);
    for (size_t k = 0; k < f.nodes.size(); ++k) {
      const Node& n = f.nodes[k];
      if (n.l < 0 || n.is_const) continue;
      cost[k] = own_cost(f, n) + cost[n.l] + cost[n.r];
      if (cost[k] == 0) continue;                        // sums of variables: a row would save nothing
      const std::pair<Cand*, bool> ins = cand_emplace(FpKey{fps[i][k].a, fps[i][k].b}, Cand{(uint32_t)i, (int32_t)k, cost[k], 1, false, -1});
      if (!ins.second) ++ins.first->refs;
      cand_of[i][k] = ins.first;
    }
  }
  // first selection: everything that occurs at least twice anywhere.  (A single function that was not cut keeps its
  // shared sub-expressions in temp slots, inside the one kernel: rows would be a second launch and more memory traffic.)
  const bool want_rows = nf >= 2 || !groups.empty();
  for (Cand& c : cands) c.selected = want_rows && c.refs >= 2;
  // from here on candidates only ever drop out: a node whose candidate is not selected now needs no second look
  for (size_t i = 0; i < nu; ++i)
    for (Cand*& c : cand_of[i])
      if (c && !c->selected) c = nullptr;
  // occurrences, top down: below a selected sub-tree nothing is counted (its inside is evaluated by the row's own
  // function).  Candidates that are left with fewer than two uses, or save less than two multiplications, are
  // dropped, which exposes what they contain -- repeat until nothing changes.
  auto count_refs = [&]() {
    for (Cand& c : cands) c.refs = 0;
    std::vector<int32_t> st;
    for (size_t i = 0; i < nu; ++i) {
      const Function& f = *unit[i];
      st.assign(1, f.root);
      while (!st.empty()) {
        const int32_t k = st.back();
        st.pop_back();
        const Node& n = f.nodes[k];
        if (n.l < 0 || n.is_const) continue;
        Cand* c = cand_of[i][k];
        if (c && c->selected) { ++c->refs; continue; }
        st.push_back(n.l);
        st.push_back(n.r);
      }
    }
  };
  for (int round = 0; round < 64; ++round) {
    count_refs();
    bool changed = false;
    for (Cand& c : cands) {
      if (c.selected && (c.refs < 2 || (uint64_t)(c.refs - 1) * c.cost < 2)) { c.selected = false; changed = true; }
    }
    if (!changed) break;
  }
  std::vector<Cand*> chosen;
  for (Cand& c : cands)
    if (c.selected) chosen.push_back(&c);
  auto benefit = [](const Cand* c) { return (uint64_t)(c->refs - 1) * c->cost; };
  std::sort(chosen.begin(), chosen.end(), [&](const Cand* x, const Cand* y) {
    if (benefit(x) != benefit(y)) return benefit(x) > benefit(y);
    return x->fn != y->fn ? x->fn < y->fn : x->node < y->node;       // deterministic
  });
  const size_t limit = table_row_limit ? table_row_limit : env_u32("FFGPU_TABLE_ROWS", 192);
  if (chosen.size() > limit) {
    for (size_t k = limit; k < chosen.size(); ++k) chosen[k]->selected = false;
    chosen.resize(limit);
    count_refs();
    std::vector<Cand*> kept;
    for (Cand* c : chosen) {
      if (c->refs >= 2) kept.push_back(c);
      else c->selected = false;
    }
    chosen.swap(kept);
  }
  if (chosen.empty() && !any_cut) return;
  // rows in the order of first appearance (deterministic tapes)
  std::sort(chosen.begin(), chosen.end(), [](const Cand* x, const Cand* y) { return x->fn != y->fn ? x->fn < y->fn : x->node < y->node; });
  std::vector<Function> table(chosen.size());
  for (size_t r = 0; r < chosen.size(); ++r) {
    chosen[r]->row = (int32_t)r;
    copy_tree(*unit[chosen[r]->fn], chosen[r]->node, table[r], [](int32_t) { return (int64_t)-1; });
    compute_needs(table[r]);
  }
  std::vector<std::vector<Fp>> tfps(table.size());
  for (size_t r = 0; r < table.size(); ++r) fingerprints(table[r], tfps[r]);
  // every unit with its selected occurrences replaced by their rows (equal fingerprints are confirmed exactly) and
  // the placeholders of the partial sums by theirs.  The partial sums (level 1) get the LOW row numbers, the level-0
  // rows the ones after them: the function that adds up its partial sums then reaches them with the 8-bit variable
  // fields of ADDVARS (eight rows per instruction instead of one -- the sum is a serial walk at small batches).
  const uint32_t R0 = (uint32_t)table.size(), R1 = (uint32_t)groups.size();
  auto lower = [&](size_t u, Function& dst) {
    const Function& f = *unit[u];
    copy_tree(f, f.root, dst, [&](int32_t k) -> int64_t {
      const Node& n = f.nodes[k];
      if (n.l < 0 || n.is_const) return -1;
      const Cand* c = cand_of[u][k];
      if (!c || !c->selected) return -1;
      const int32_t r = c->row;
      if (!same_structure(f, fps[u], k, table[r], tfps[r], table[r].root)) return -1;
      return (int64_t)n_base + R1 + r;
    });
    for (Node& n : dst.nodes)
      if ((n.kind == N_VAR || n.kind == N_NEGVAR) && n.var >= GROUP_BASE) n.var = n_base + (n.var - GROUP_BASE);
    compute_needs(dst);
  };
  lowered_.resize(nf);
  std::vector<Function> lowered_groups(R1);
  for (size_t u = 0; u < nu; ++u) {
    if (unit_fn[u] >= 0) {
      lower(u, lowered_[unit_fn[u]]);       // (the RPN strings stay with functions[]: nothing reads them from a lowered copy)
    } else {
      lower(u, lowered_groups[(size_t)(-1 - unit_fn[u])]);
    }
  }
  for (size_t i = 0; i < nf; ++i)
    if (dup_of[i] >= 0) lowered_[i] = functions[i];
  // the rows' own collections: level 0 over the variables, level 1 over the variables and the level-0 rows
  auto add_level = [&](std::vector<Function>&& fns, uint32_t extra_vars) -> bool {
    pre.emplace_back();
    Collection& pc = pre.back();
    pc.vars = vars;
    pc.var_index = var_index;
    for (uint32_t r = 0; r < extra_vars; ++r) pc.vars.push_back("$row" + std::to_string(r));
    pc.n_input_vars = n_base;
    pc.check_is_equal = false;
    pc.functions = std::move(fns);
    pc.enable_fuse = enable_fuse;
    pc.enable_cse = enable_cse;
    pc.enable_cross = enable_cross;
    pc.enable_table = false;
    pc.temp_budget = temp_budget;
    pc.small_prime_mode_ = small_prime_mode_;
    std::string perr;
    return pc.finish(perr);
  };
  bool ok = true;
  pre_row_offset_.clear();
  if (R0) { pre_row_offset_.push_back(R1); ok = add_level(std::move(table), 0); }
  if (ok && R1) { pre_row_offset_.push_back(0); ok = add_level(std::move(lowered_groups), R1 + R0); }
  if (!ok) {                       // never expected; the functions are then compiled as written
    pre.clear();
    pre_row_offset_.clear();
    lowered_.clear();
    return;
  }
  n_table_rows = R0 + R1;
  n_split_rows = R1;
}

bool Collection::add_function(std::string text, uint64_t index, bool strip_blanks, std::string& err, bool validate,
                              bool dedupe, bool no_duplicates, uint64_t* position) {
  if (strip_blanks) {
    text.erase(std::remove_if(text.begin(), text.end(), [](char c) { return c == ' ' || c == '\t' || c == '\n'; }),
               text.end());
    if (text.empty()) return true;          // blank chunk between ';' adds no function
  }
  if (validate && !validate_expr(text, index, err)) return false;
  Function f;
  if (!parse_expr(text, f.rpn, err)) return false;
  if (!build_tree(f, err)) return false;
  if (check_is_equal && dedupe) {
    // fingerprint = values at one random point in each of two primes (reference: :125-140)
    std::vector<uint64_t> v1, v2;
    eval_nodes(f, P1, chk_pts1_.data(), v1);
    eval_nodes(f, P0, chk_pts2_.data(), v2);
    const std::string key = std::to_string(v1[f.root]) + ":" + std::to_string(v2[f.root]);
    auto it = chk_map_.find(key);
    if (it != chk_map_.end()) {
      if (!no_duplicates) evaluation_positions.push_back(it->second);
      if (position) *position = it->second;
      return true;
    }
    chk_map_.emplace(key, (uint64_t)functions.size());
    evaluation_positions.push_back((uint64_t)functions.size());
  }
  if (position) *position = (uint64_t)functions.size();
  functions.push_back(std::move(f));
  return true;
}

bool Collection::finish(std::string& err) {
  (void)err;
  ++rebuilds;
  stale = false;                 // resolve() below must not come back here
  n_out = check_is_equal ? evaluation_positions.size() : functions.size();
  // output slots per distinct function
  std::vector<std::vector<uint32_t>> slots(functions.size());
  if (check_is_equal) {
    for (size_t k = 0; k < evaluation_positions.size(); ++k) slots[evaluation_positions[k]].push_back((uint32_t)k);
  } else {
    for (size_t k = 0; k < functions.size(); ++k) slots[k].push_back((uint32_t)k);
  }
  // Functions of equal structure are evaluated once and stored to all their outputs -- check_is_equal without
  // the flag for what is equal as written (the flag compares VALUES at a random point, reference :125-140, and
  // also finds functions that are equal only algebraically).
  std::vector<int32_t> dup_of(functions.size(), -1);
  n_structural_duplicates = 0;
  std::vector<std::vector<Fp>> fps;
  if (enable_cse && enable_cross && functions.size() > 1) {
    fps.resize(functions.size());
    std::unordered_map<uint64_t, std::vector<uint32_t>> seen;
    for (size_t i = 0; i < functions.size(); ++i) {
      fingerprints(functions[i], fps[i]);
      const Fp h = fps[i][functions[i].root];
      auto& bucket = seen[h.a ^ mix64(h.b)];
      for (uint32_t j : bucket)
        if (same_structure(functions[i], fps[i], functions[i].root, functions[j], fps[j], functions[j].root)) { dup_of[i] = (int32_t)j; break; }
      if (dup_of[i] < 0) { bucket.push_back((uint32_t)i); continue; }
      auto& to = slots[dup_of[i]];
      to.insert(to.end(), slots[i].begin(), slots[i].end());
      slots[i].clear();
      ++n_structural_duplicates;
    }
  }
  // sub-expressions shared by several functions become extra variable rows (build_table); from here on the
  // functions are the lowered ones (efn)
  build_table(dup_of);
  if (!lowered_.empty()) {
    fps.resize(functions.size());
    for (size_t i = 0; i < functions.size(); ++i) fingerprints(efn(i), fps[i]);
  }
  // Evaluation order: functions over the same denominator are brought together (the group takes the place of its
  // first member), so that a denominator that repeats anywhere in the collection is inverted once.  Outputs are
  // addressed by their slot, the order of evaluation is free.  fn_start is indexed by POSITION in this order.
  order_.resize(functions.size());
  for (size_t i = 0; i < functions.size(); ++i) order_[i] = (uint32_t)i;
  if (!fps.empty()) {
    std::unordered_map<uint64_t, std::vector<uint32_t>> groups;
    std::vector<uint64_t> key(functions.size(), 0);
    for (size_t i = 0; i < functions.size(); ++i) {
      const Function& f = efn(i);
      const Node& rt = f.nodes[f.root];
      if (dup_of[i] >= 0 || rt.kind != N_DIV || rt.is_const || f.nodes[rt.r].is_const) continue;
      const Fp h = fps[i][rt.r];
      key[i] = (h.a ^ mix64(h.b)) | 1;
      groups[key[i]].push_back((uint32_t)i);
    }
    std::vector<char> placed(functions.size(), 0);
    size_t pos = 0;
    for (size_t i = 0; i < functions.size(); ++i) {
      if (placed[i]) continue;
      if (!key[i]) { order_[pos++] = (uint32_t)i; placed[i] = 1; continue; }
      for (uint32_t j : groups[key[i]]) { order_[pos++] = j; placed[j] = 1; }
    }
  }
  pos_of_.resize(functions.size());
  for (size_t k = 0; k < order_.size(); ++k) pos_of_[order_[k]] = (uint32_t)k;
  no_cse_fn_.assign(functions.size(), 0);
  cse_fallbacks = 0;
  for (int pass = 0; pass < 2; ++pass) {
    tape.clear();
    patches.clear();
    stack_depth = 0;
    n_temps = 0;
    n_slots = 0;
    work = WorkCount{};
    fn_start.clear();
    // Shared denominators: a run of consecutive functions N_i / D over the same D (equal structure) inverts D once.
    // The first function of the run parks 1/D in temp slot 0 (den_mode 1), the others multiply their numerator
    // by it (den_mode 2) -- no denominator code, no inversion, no barrier.  a/0 = 0 is kept: 1/0 = 0 on the
    // device (reference FFInt::operator/=).  A chunk of the tape never starts inside a run (chunk tables below).
    den_mode_.assign(functions.size(), 0);
    n_shared_denominators = 0;
    if (enable_cse && !fps.empty()) {
      auto divisor = [&](size_t i) -> int32_t {
        const Function& f = efn(i);
        const Node& rt = f.nodes[f.root];
        if (no_cse_fn_[i] || rt.kind != N_DIV || rt.is_const || f.nodes[rt.r].is_const) return -1;
        return rt.r;
      };
      int64_t head = -1;                        // first function of the current run
      for (size_t k = 0; k < functions.size(); ++k) {
        const size_t i = order_[k];
        if (dup_of[i] >= 0) continue;           // emits nothing: transparent
        const int32_t d = divisor(i);
        bool joins = false;
        if (head >= 0 && d >= 0) {
          const int32_t dh = divisor((size_t)head);
          joins = same_structure(efn(i), fps[i], d, efn(head), fps[head], dh);
        }
        if (joins) {
          den_mode_[head] = 1;
          den_mode_[i] = 2;
          ++n_shared_denominators;
        } else {
          head = d >= 0 ? (int64_t)i : -1;
        }
      }
    }
    for (size_t k = 0; k < functions.size(); ++k) {
      const size_t i = order_[k];
      fn_start.push_back((uint32_t)tape.size());
      if (dup_of[i] < 0) emit_function((uint32_t)i, slots[i], den_mode_[i]);   // a duplicate's outputs are stored by its twin
      count_work(functions[i]);
    }
    fn_start.push_back((uint32_t)tape.size());
    peephole();
    if (enable_fuse) fuse();
    // Translation check of every function at one pseudo-random point: the tape must reproduce the value of
    // the expression tree.  A function that fails is re-emitted as a plain tree (no sharing) on the second
    // pass; if that fails too the compilation is refused -- never a wrong probe value.
    if (!resolve_constants(P0, err)) return false;
    std::vector<uint64_t> pt(vars.size());
    uint64_t seed = 0x243F6A8885A308D3ULL;
    for (auto& v : pt) v = 1 + splitmix64(seed) % (P0 - 1);
    // the shared rows at that point, level by level, from their own expression trees (each level's tape has been
    // checked against them by its own finish()); a level sees the variables and the rows of the levels before it
    pt.resize(vars.size() + n_table_rows, 0);
    for (size_t l = 0; l < pre.size(); ++l) {
      std::vector<uint64_t> val;
      for (size_t r = 0; r < pre[l].functions.size(); ++r) {
        eval_nodes(pre[l].functions[r], P0, pt.data(), val);
        pt[vars.size() + pre_row_offset_[l] + r] = val[pre[l].functions[r].root];
      }
    }
    bool again = false;
    uint64_t chk_temps[16] = {};
    bool chk_temp_set[16] = {};
    for (size_t k = 0; k < functions.size(); ++k) {
      const size_t i = order_[k];
      if (dup_of[i] >= 0 || check_function((uint32_t)i, P0, pt.data(), chk_temps, chk_temp_set)) continue;
      if (pass == 1 || no_cse_fn_[i] || !enable_cse) {
        err = "internal error: the tape of function " + std::to_string(i) + " does not reproduce its expression";
        return false;
      }
      no_cse_fn_[i] = 1;
      ++cse_fallbacks;
      again = true;
    }
    if (!again) break;
  }
  // work the tape executes per probe (the imm fields are resolved for P0 here; exponents of literal powers
  // do not depend on the prime)
  executed = WorkCount{};
  {
    auto chain = [](uint64_t e) -> uint64_t {
      return e <= 1 ? 0 : (uint64_t)(63 - __builtin_clzll(e)) + (uint64_t)__builtin_popcountll(e) - 1;
    };
    for (const Instr& in : tape) {
      ++executed.tokens;
      if (in.src == SRC_VAR || in.src == SRC_NEGVAR) ++executed.max_rpn_depth;   // reused: variable row reads
      switch (in.op) {
        case OP_ADD: case OP_SUB: case OP_NEG: case OP_ACCUM: ++executed.ma; break;
        case OP_MUL: case OP_SQR: ++executed.mm; break;
        case OP_MULADD: ++executed.mm; ++executed.ma; break;
        case OP_INV: ++executed.inv; break;
        case OP_POWU:
          if (in.imm < (P0 >> 1)) executed.mm += chain(in.imm);
          else { ++executed.inv; executed.mm += chain(P0 - in.imm - 1); }
          break;
        case OP_NPOWU: {
          const int32_t e = ffhost::to_neg_int(in.imm);
          if (e >= 0) executed.mm += chain((uint64_t)e);
          else { ++executed.inv; executed.mm += chain((uint64_t)(-(int64_t)e)); }
          break;
        }
        case OP_POWL: case OP_NPOWL: executed.mm += 94; break;
        case OP_TERM: {
          const uint8_t mode = (in.flags >> TM_SHIFT) & 7;
          executed.mm += chain(in.arg >> 16) + ((in.flags & F_CMUL) ? 1 : 0) + (mode == TM_MUL ? 1 : 0);
          if (mode == TM_ADD || mode == TM_SUB || mode == TM_RSUB) ++executed.ma;
          break;
        }
        case OP_ACCUM2: executed.ma += 2; break;
        case OP_ADDVARS: executed.ma += in.arg & 0xff; executed.max_rpn_depth += (in.arg & 0xff) - 1; break;
        case OP_MULVARS: executed.mm += in.arg & 0xff; executed.max_rpn_depth += (in.arg & 0xff) - 1; break;
        default: break;
      }
      if (in.op < OP_TERM && (in.flags & F_NEGR)) ++executed.ma;
    }
  }
  for (const Collection& pc : pre) {
    const WorkCount& w = pc.executed;
    executed.tokens += w.tokens;
    executed.mm += w.mm;
    executed.ma += w.ma;
    executed.inv += w.inv;
    executed.max_rpn_depth += w.max_rpn_depth;
  }
  // chunk tables for C = 1, 2, 4, ... while a chunk still holds a function
  chunk_table.clear();
  chunk_table_off.clear();
  max_chunks = 1;
  const uint32_t total = (uint32_t)tape.size();
  for (unsigned C = 1; C <= 1024 && C <= std::max<size_t>(functions.size(), 1); C <<= 1) {
    chunk_table_off.push_back((uint32_t)chunk_table.size());
    size_t f = 0;
    for (unsigned k = 0; k < C; ++k) {
      const uint64_t target = (uint64_t)total * k / C;
      while (f + 1 < fn_start.size() && fn_start[f] < target) ++f;
      // a chunk starts with temp slot 0 undefined: back up to the function that fills it
      size_t g = f;
      while (g > 0 && g < functions.size() && (den_mode_[order_[g]] == 2 || dup_of[order_[g]] >= 0)) --g;
      chunk_table.push_back(fn_start[g]);
    }
    chunk_table.push_back(total);
    max_chunks = C;
  }
  // One more table with ONE function per chunk (as far as functions may be separated): a collection of a few heavy
  // functions -- the partial sums of a long sum -- evaluated at a handful of points wants every function on an SM of
  // its own, and the power-of-two tables above put two of 194 into every second chunk.
  exact_chunks = 0;
  exact_table = 0;
  if (functions.size() > max_chunks && functions.size() <= 1024) {
    std::vector<uint32_t> starts;
    for (size_t g = 0; g < functions.size(); ++g) {
      if (den_mode_[order_[g]] == 2 || dup_of[order_[g]] >= 0) continue;            // stays with the function before it
      if (!starts.empty() && starts.back() == fn_start[g]) continue;
      starts.push_back(fn_start[g]);
    }
    if (starts.size() > max_chunks && !starts.empty() && starts[0] == 0) {
      exact_table = (uint32_t)chunk_table_off.size();
      chunk_table_off.push_back((uint32_t)chunk_table.size());
      chunk_table.insert(chunk_table.end(), starts.begin(), starts.end());
      chunk_table.push_back(total);
      exact_chunks = (uint32_t)starts.size();
    }
  }
  Instr halt{};
  halt.op = OP_HALT;
  tape.push_back(halt);
  tape.push_back(halt);            // the interpreter prefetches one instruction ahead
  prime = 0;
  stale = false;
  return true;
}

// Strength reduction on the finished tape and assignment of the dense dispatch codes.
void Collection::peephole() {
  std::vector<int32_t> patch_of(tape.size(), -1);
  for (size_t i = 0; i < patches.size(); ++i) patch_of[patches[i].instr] = (int32_t)i;
  for (size_t i = 0; i < tape.size(); ++i) {
    Instr& in = tape[i];
    // acc - c  ->  acc + (-c);   acc + (-v) -> acc - v;   acc - (-v) -> acc + v
    if (in.op == OP_SUB && in.src == SRC_CONST && !(in.flags & F_SWAP) && patch_of[i] >= 0) {
      in.op = OP_ADD;
      patches[patch_of[i]].negate = !patches[patch_of[i]].negate;
    } else if (in.op == OP_ADD && in.src == SRC_NEGVAR) {
      in.op = OP_SUB;
      in.src = SRC_VAR;
    } else if (in.op == OP_SUB && in.src == SRC_NEGVAR && !(in.flags & F_SWAP)) {
      in.op = OP_ADD;
      in.src = SRC_VAR;
    }
    uint8_t fx = FX_GENERIC;
    const bool swap = in.flags & F_SWAP, push = in.flags & F_PUSH, init = in.flags & F_INIT;
    const uint8_t other = in.flags & ~(F_SWAP | F_PUSH | F_INIT);
    if (!other) {
      switch (in.op) {
        case OP_LOAD:
          if (in.src == SRC_VAR) fx = push ? FX_PUSH_LOAD_VAR : FX_LOAD_VAR;
          else if (in.src == SRC_CONST) fx = push ? FX_PUSH_LOAD_CONST : FX_LOAD_CONST;
          else if (in.src == SRC_TMP) fx = push ? FX_PUSH_LOAD_TMP : FX_LOAD_TMP;
          break;
        case OP_ADD:
          fx = in.src == SRC_VAR ? FX_ADD_VAR : in.src == SRC_CONST ? FX_ADD_CONST : in.src == SRC_POP ? FX_ADD_POP :
               in.src == SRC_TMP ? FX_ADD_TMP : FX_GENERIC;
          break;
        case OP_SAVE: fx = FX_SAVE; break;
        case OP_ACCUM:
          if (init) fx = swap ? FX_GENERIC : FX_ACCUM_INIT;
          else fx = swap ? FX_ACCUM_SUB : FX_ACCUM_ADD;
          break;
        case OP_SUB:
          if (in.src == SRC_VAR) fx = swap ? FX_RSUB_VAR : FX_SUB_VAR;
          else if (in.src == SRC_CONST && swap) fx = FX_RSUB_CONST;
          else if (in.src == SRC_POP) fx = swap ? FX_RSUB_POP : FX_SUB_POP;
          break;
        case OP_MUL:
          fx = in.src == SRC_VAR ? FX_MUL_VAR : in.src == SRC_CONST ? FX_MUL_CONST : in.src == SRC_POP ? FX_MUL_POP :
               in.src == SRC_TMP ? FX_MUL_TMP : FX_GENERIC;
          break;
        case OP_POWU: fx = FX_POWU; break;
        case OP_INV: fx = FX_INV; break;
        case OP_STORE: fx = FX_STORE; break;
        case OP_SQR: fx = FX_SQR; break;
        default: break;
      }
    }
    in.fast = fx;
  }
}

// Fusion of frequent instruction runs into one instruction each (one dispatch of the interpreter instead
// of two to four; the term forms also keep the old ACC out of the explicit spill protocol).  Runs never cross
// a function boundary (every function ends in a STORE, which is never part of a run).  Purely a re-encoding:
// the operations and their order are unchanged.
void Collection::fuse() {
  const size_t n = tape.size();
  std::vector<int32_t> patch_of(n, -1);
  for (size_t i = 0; i < patches.size(); ++i) patch_of[patches[i].instr] = (int32_t)i;
  std::vector<Instr> out;
  out.reserve(n);
  std::vector<uint32_t> new_index(n + 1, 0);
  std::vector<char> patch_dead(patches.size(), 0);

  // value of a literal exponent if it is the same small number for every prime
  auto small_literal = [&](int32_t pi, uint32_t& e) -> bool {
    if (pi < 0 || small_prime_mode_) return false;      // tiny primes: exponents stay per-prime residues (resolve())
    const ConstPatch& cp = patches[pi];
    if (cp.inverse || cp.negate) return false;
    const Function& f = efn(cp.fn);
    const Node& nd = f.nodes[cp.node];
    if (nd.kind != N_NUM) return false;
    const std::string& t = f.literals[nd.lit].text;
    if (t.empty() || t.size() > 6) return false;
    for (char ch : t)
      if (ch < '0' || ch > '9') return false;
    e = (uint32_t)leading_digits(t, 0);
    return e >= 1 && e < 65536;
  };
  auto plain = [](const Instr& in, uint8_t op, uint8_t src) { return in.op == op && in.src == src && in.flags == 0; };

  size_t i = 0;
  while (i < n) {
    const Instr& a = tape[i];
    const size_t first = i;
    uint32_t e = 0;
    // ---- LOAD.VAR[.P], POWU e, [MUL.CONST], [ADD/SUB/MUL.POP]  ->  TERM
    if (a.op == OP_LOAD && a.src == SRC_VAR && (a.flags & ~F_PUSH) == 0 && a.arg < 65536 && i + 1 < n &&
        plain(tape[i + 1], OP_POWU, SRC_CONST) && small_literal(patch_of[i + 1], e)) {
      Instr t{};
      t.op = OP_TERM;
      t.src = SRC_VAR;
      t.arg = a.arg | (e << 16);
      t.flags = a.flags & F_PUSH;
      t.fast = FX_TERM;
      patch_dead[patch_of[i + 1]] = 1;                       // the exponent now lives in arg
      size_t j = i + 2;
      int32_t moved_patch = -1;
      if (j < n && plain(tape[j], OP_MUL, SRC_CONST) && patch_of[j] >= 0) {
        t.flags |= F_CMUL;
        moved_patch = patch_of[j];
        ++j;
      }
      if ((t.flags & F_PUSH) && j < n && tape[j].src == SRC_POP && (tape[j].flags & ~F_SWAP) == 0 &&
          (tape[j].op == OP_ADD || tape[j].op == OP_SUB || tape[j].op == OP_MUL)) {
        // ACC holds the term T, the popped value is the old accumulator A
        uint8_t mode = TM_ADD;
        if (tape[j].op == OP_MUL) mode = TM_MUL;
        else if (tape[j].op == OP_SUB) mode = (tape[j].flags & F_SWAP) ? TM_SUB : TM_RSUB;   // swap: A - T; else T - A
        t.flags = (uint8_t)((t.flags & ~F_PUSH) | (mode << TM_SHIFT));
        ++j;
      }
      for (size_t k = first; k < j; ++k) new_index[k] = (uint32_t)out.size();
      if (moved_patch >= 0) patches[moved_patch].instr = (uint32_t)out.size();
      out.push_back(t);
      i = j;
      continue;
    }
    // ---- LOAD.VAR.P, MUL.CONST, ADD/SUB/MUL.POP  ->  TERM with exponent 1 (c * x as a term of a sum; with shared
    //      rows the powers of a polynomial's monomials arrive as plain variables)
    if (a.op == OP_LOAD && a.src == SRC_VAR && a.flags == F_PUSH && a.arg < 65536 && i + 2 < n &&
        plain(tape[i + 1], OP_MUL, SRC_CONST) && patch_of[i + 1] >= 0 && tape[i + 2].src == SRC_POP &&
        (tape[i + 2].flags & ~F_SWAP) == 0 &&
        (tape[i + 2].op == OP_ADD || tape[i + 2].op == OP_SUB || tape[i + 2].op == OP_MUL)) {
      Instr t{};
      t.op = OP_TERM;
      t.src = SRC_VAR;
      t.arg = a.arg | (1u << 16);
      t.fast = FX_TERM;
      uint8_t mode = TM_ADD;
      if (tape[i + 2].op == OP_MUL) mode = TM_MUL;
      else if (tape[i + 2].op == OP_SUB) mode = (tape[i + 2].flags & F_SWAP) ? TM_SUB : TM_RSUB;
      t.flags = (uint8_t)(F_CMUL | (mode << TM_SHIFT));
      for (size_t k = i; k < i + 3; ++k) new_index[k] = (uint32_t)out.size();
      patches[patch_of[i + 1]].instr = (uint32_t)out.size();
      out.push_back(t);
      i += 3;
      continue;
    }
    // ---- LOAD.VAR[.P], MUL.CONST  ->  TERM with exponent 1 in load mode (the innermost c * x of a Horner nest)
    if (a.op == OP_LOAD && a.src == SRC_VAR && (a.flags & ~F_PUSH) == 0 && a.arg < 65536 && i + 1 < n &&
        plain(tape[i + 1], OP_MUL, SRC_CONST) && patch_of[i + 1] >= 0) {
      Instr t{};
      t.op = OP_TERM;
      t.src = SRC_VAR;
      t.arg = a.arg | (1u << 16);
      t.fast = FX_TERM;
      t.flags = (uint8_t)((a.flags & F_PUSH) | F_CMUL | (TM_LOAD << TM_SHIFT));
      new_index[i] = new_index[i + 1] = (uint32_t)out.size();
      patches[patch_of[i + 1]].instr = (uint32_t)out.size();
      out.push_back(t);
      i += 2;
      continue;
    }
    // ---- ACCUM, ACCUM (two sums fed by the same term)  ->  ACCUM2
    if (a.op == OP_ACCUM && i + 1 < n && tape[i + 1].op == OP_ACCUM && a.arg < 256 && tape[i + 1].arg < 256 &&
        a.arg != tape[i + 1].arg && (a.flags & ~(F_SWAP | F_INIT)) == 0 && (tape[i + 1].flags & ~(F_SWAP | F_INIT)) == 0) {
      const Instr& b = tape[i + 1];
      Instr t{};
      t.op = OP_ACCUM2;
      t.arg = a.arg | (b.arg << 8);
      t.flags = (uint8_t)((a.flags & (F_SWAP | F_INIT)) | ((b.flags & F_SWAP) ? F_SWAP2 : 0) | ((b.flags & F_INIT) ? F_INIT2 : 0));
      t.fast = FX_ACCUM2;
      new_index[i] = new_index[i + 1] = (uint32_t)out.size();
      out.push_back(t);
      i += 2;
      continue;
    }
    // ---- runs of ADD.VAR / SUB.VAR  ->  ADDVARS;  runs of MUL.VAR  ->  MULVARS
    auto addlike = [&](const Instr& in) { return (in.op == OP_ADD || in.op == OP_SUB) && in.src == SRC_VAR && in.flags == 0 && in.arg < 256; };
    auto mullike = [&](const Instr& in) { return in.op == OP_MUL && in.src == SRC_VAR && in.flags == 0 && in.arg < 256; };
    const bool is_add = addlike(a), is_mul = mullike(a);
    // a constant added after a multiplication by a variable is a Horner step (the form the FireFly paper
    // recommends for large inputs): MUL.VAR, ADD.CONST | (c - ACC)  ->  MULADD
    auto horner_tail = [&](size_t k) {
      return k < n && tape[k].src == SRC_CONST && patch_of[k] >= 0 &&
             ((tape[k].op == OP_ADD && tape[k].flags == 0) || (tape[k].op == OP_SUB && tape[k].flags == F_SWAP));
    };
    if (is_add || is_mul) {
      size_t j = i;
      uint64_t ids = 0;
      uint32_t signs = 0, cnt = 0;
      while (j < n && cnt < 8 && (is_add ? addlike(tape[j]) : mullike(tape[j]))) {
        ids |= (uint64_t)tape[j].arg << (8 * cnt);
        if (tape[j].op == OP_SUB) signs |= 1u << cnt;
        ++cnt;
        ++j;
      }
      if (is_mul && horner_tail(j)) {      // the last multiplication of the run belongs to the Horner step
        --cnt;
        --j;
      }
      if (is_mul && cnt == 0) {
        Instr t{};
        t.op = OP_MULADD;
        t.src = SRC_VAR;
        t.arg = tape[j].arg;
        t.flags = tape[j + 1].op == OP_SUB ? F_SWAP : 0;
        t.fast = FX_MULADD;
        new_index[j] = new_index[j + 1] = (uint32_t)out.size();
        patches[patch_of[j + 1]].instr = (uint32_t)out.size();
        out.push_back(t);
        i = j + 2;
        continue;
      }
      if (cnt >= 2) {
        Instr t{};
        t.op = is_add ? OP_ADDVARS : OP_MULVARS;
        t.src = SRC_VAR;
        t.arg = cnt | (signs << 8);
        t.imm = ids;
        t.fast = is_add ? FX_ADDVARS : FX_MULVARS;
        for (size_t k = i; k < j; ++k) new_index[k] = (uint32_t)out.size();
        out.push_back(t);
        i = j;
        continue;
      }
    }
    new_index[i] = (uint32_t)out.size();
    if (patch_of[i] >= 0) patches[patch_of[i]].instr = (uint32_t)out.size();
    out.push_back(a);
    ++i;
  }
  new_index[n] = (uint32_t)out.size();
  for (uint32_t& s0 : fn_start) s0 = new_index[s0];
  std::vector<ConstPatch> kept;
  kept.reserve(patches.size());
  for (size_t k = 0; k < patches.size(); ++k)
    if (!patch_dead[k]) kept.push_back(patches[k]);
  patches.swap(kept);
  tape.swap(out);
}

static void init_vars(Collection& c, const std::vector<std::string>& vars, bool chk) {
  c.vars = vars;
  c.var_index.clear();
  for (uint32_t i = 0; i < vars.size(); ++i) c.var_index.emplace(vars[i], i);   // first declaration wins
  c.check_is_equal = chk;
  c.functions.clear();
  c.evaluation_positions.clear();
}

void Collection::init_check_points() {
  chk_map_.clear();
  base_functions_.clear();
  if (!check_is_equal) return;
  uint64_t s = 0x4d595df4d0f33173ULL;
  chk_pts1_.resize(vars.size());
  chk_pts2_.resize(vars.size());
  for (auto& v : chk_pts1_) v = 1 + splitmix64(s) % (P1 - 1);
  for (auto& v : chk_pts2_) v = 1 + splitmix64(s) % (P0 - 1);
}

bool Collection::compile_strings(const std::vector<std::string>& funs, const std::vector<std::string>& vars_,
                                 bool chk, std::string& err) {
  init_vars(*this, vars_, chk);
  init_check_points();
  for (size_t i = 0; i < funs.size(); ++i)
    if (!add_function(funs[i], i, false, err)) return false;
  return finish(err);
}

bool Collection::compile_text(const std::string& text, const std::vector<std::string>& vars_, bool chk,
                              std::string& err) {
  init_vars(*this, vars_, chk);
  init_check_points();
  size_t start = 0;
  uint64_t line = 1;
  while (start <= text.size()) {
    size_t end = text.find(';', start);
    if (end == std::string::npos) end = text.size();
    if (end > start || end < text.size()) {
      if (!add_function(text.substr(start, end - start), line, true, err)) return false;
    }
    ++line;
    if (end == text.size()) break;
    start = end + 1;
  }
  return finish(err);
}

// Scanned factors (reference: Reconstructor::get_job divides output i of the black box by
// parsed_factors[i].evaluate_pre(values)[0], Reconstructor.hpp:3392-3396,3449-3457; MPIWorker.hpp:486-495).
// Here the division joins the function's expression tree -- out_i = f_i / g_i with FFInt's a/0 = 0 -- so it runs
// on the device inside the same tape.  The factor expressions have their own variable names (FireFly writes
// them over x1..xn, ParserUtils.cpp:182-185), matched to the collection's variables by position.
bool Collection::apply_factors(const std::vector<std::pair<uint32_t, std::string>>& factors,
                               const std::vector<std::string>& factor_vars, std::string& err) {
  if (factor_vars.size() != vars.size()) {
    err = "the factors must be written over as many variables as the collection has";
    return false;
  }
  if (stale) n_out = check_is_equal ? evaluation_positions.size() : functions.size();
  if (base_functions_.empty() && !functions.empty()) {
    // one function per OUTPUT from here on: duplicates may get different factors
    for (size_t k = 0; k < n_out; ++k) base_functions_.push_back(functions[check_is_equal ? evaluation_positions[k] : k]);
  }
  std::vector<Function> nf = base_functions_;
  std::unordered_map<std::string, uint32_t> fidx, saved;
  for (uint32_t i = 0; i < factor_vars.size(); ++i) fidx.emplace(factor_vars[i], i);
  for (const auto& fc : factors) {
    if (fc.first >= nf.size()) {
      err = "factor for output " + std::to_string(fc.first) + ": the collection returns " + std::to_string(nf.size()) + " values";
      return false;
    }
    std::string text = fc.second;
    text.erase(std::remove_if(text.begin(), text.end(), [](char c) { return c == ' ' || c == '\t' || c == '\n'; }), text.end());
    Function g;
    saved.swap(var_index);
    var_index = fidx;
    const bool ok = validate_expr(text, fc.first, err) && parse_expr(text, g.rpn, err) && build_tree(g, err);
    var_index.swap(saved);
    saved.clear();
    if (!ok) return false;
    Function& f = nf[fc.first];
    const int32_t off = (int32_t)f.nodes.size();
    const uint32_t lit_off = (uint32_t)f.literals.size();
    for (Node nd : g.nodes) {
      if (nd.l >= 0) nd.l += off;
      if (nd.r >= 0) nd.r += off;
      if (nd.kind == N_NUM) nd.lit += lit_off;
      f.nodes.push_back(nd);
    }
    f.literals.insert(f.literals.end(), g.literals.begin(), g.literals.end());
    Node div{};
    div.kind = N_DIV;
    div.l = f.root;
    div.r = g.root + off;
    div.is_const = f.nodes[div.l].is_const && f.nodes[div.r].is_const;
    f.nodes.push_back(div);
    f.root = (int32_t)f.nodes.size() - 1;
    f.rpn.insert(f.rpn.end(), g.rpn.begin(), g.rpn.end());
    f.rpn.emplace_back("/");
    compute_needs(f);
  }
  functions.swap(nf);
  check_is_equal = false;
  evaluation_positions.clear();
  return finish(err);
}

// ------------------------------------------------------------------------------------------
// 9. changing a compiled collection: add_otf, parse_function, add_otf_precompute, trim
// ------------------------------------------------------------------------------------------
bool Collection::append(std::string text, bool validate, bool dedupe, bool no_duplicates, uint64_t& position,
                        std::string& err) {
  if (!base_functions_.empty()) {
    err = "functions cannot be added after scanned factors were set";
    return false;
  }
  if (check_is_equal && !dedupe) {
    // parse_function on a check_is_equal parser leaves the function without an evaluation position: the
    // reference computes it and never returns its value (ShuntingYardParser.hpp:462-478)
    err = "parse_function on a parser that checks for equal functions: use add_otf";
    return false;
  }
  const size_t before = functions.size(), outs = evaluation_positions.size();
  if (!add_function(std::move(text), before, false, err, validate, dedupe, no_duplicates, &position)) return false;
  if (functions.size() != before || evaluation_positions.size() != outs) stale = true;
  n_out = check_is_equal ? evaluation_positions.size() : functions.size();
  return true;
}

bool Collection::append_rpn(const std::vector<std::string>& rpn, uint64_t& position, std::string& err) {
  if (!base_functions_.empty()) {
    err = "functions cannot be added after scanned factors were set";
    return false;
  }
  if (check_is_equal) {
    err = "add_otf_precompute on a parser that checks for equal functions: the function would have no evaluation position";
    return false;
  }
  Function f;
  f.rpn = rpn;
  if (!build_tree(f, err)) return false;
  position = (uint64_t)functions.size();
  functions.push_back(std::move(f));
  n_out = functions.size();
  stale = true;
  return true;
}

bool Collection::trim(const std::vector<uint64_t>& keep, std::vector<uint64_t>& new_pos, std::string& err) {
  if (!base_functions_.empty()) {
    err = "a collection with scanned factors cannot be trimmed";
    return false;
  }
  new_pos.assign(keep.size(), ~0ULL);
  std::unordered_map<uint64_t, size_t> want;                  // old position -> index into keep
  for (size_t k = 0; k < keep.size(); ++k) want.emplace(keep[k], k);
  std::vector<Function> nf;
  if (!check_is_equal) {
    for (size_t i = 0; i < functions.size(); ++i) {
      auto it = want.find(i);
      if (it == want.end()) continue;
      nf.push_back(std::move(functions[i]));
      new_pos[it->second] = nf.size() - 1;
    }
  } else {
    // positions are OUTPUT positions; a function that several kept outputs share is copied once
    std::vector<uint64_t> npos;
    std::unordered_map<uint64_t, uint64_t> copied;             // old function -> new function
    for (size_t i = 0; i < evaluation_positions.size(); ++i) {
      auto it = want.find(i);
      if (it == want.end()) continue;
      const uint64_t fn = evaluation_positions[i];
      auto c = copied.find(fn);
      if (c == copied.end()) {
        nf.push_back(std::move(functions[fn]));
        c = copied.emplace(fn, nf.size() - 1).first;
      }
      npos.push_back(c->second);
      new_pos[it->second] = npos.size() - 1;
    }
    evaluation_positions.swap(npos);
    // the fingerprint map goes with the old numbering (the reference leaves its check_map as it is: a later
    // add_otf of a trimmed function would return a stale position there); here later additions are compared
    // with what was kept
    for (auto it = chk_map_.begin(); it != chk_map_.end();) {
      auto c = copied.find(it->second);
      if (c == copied.end()) it = chk_map_.erase(it);
      else { it->second = c->second; ++it; }
    }
  }
  // duplicates in `keep` name the same element
  for (size_t k = 0; k < keep.size(); ++k) new_pos[k] = new_pos[want[keep[k]]];
  functions.swap(nf);
  n_out = check_is_equal ? evaluation_positions.size() : functions.size();
  stale = true;
  return true;
}

bool Collection::resolve(uint64_t p, std::string& err) {
  if (p < 3 || (p >> 63)) {
    err = "prime must be below 2^63";
    return false;
  }
  // A literal exponent is a residue like every other number of the input: the reference reduces it mod p and takes
  // exponents from p/2 up as negative powers (FFInt::pow(const FFInt&), FFInt.cpp:155-179).  For the primes FireFly works
  // with that never shows (a literal exponent is far below 2^62), and the passes that carry exponents as plain integers
  // -- TERM's exponent field, the Horner nests, which split x^e into x^m * x^(e-m) -- rely on it.  Below 2^21 they are
  // switched off: every power then goes through OP_POWU with its per-prime residue.
  const bool small = p < (1ULL << 21);
  if (small != small_prime_mode_) {
    small_prime_mode_ = small;
    stale = true;
  }
  if (stale && !finish(err)) {
    stale = true;              // nothing usable was built: evaluation keeps refusing
    prime = 0;
    return false;
  }
  return resolve_constants(p, err);
}

bool Collection::resolve_constants(uint64_t p, std::string& err) {
  std::vector<uint64_t> val;
  uint32_t cur = ~0u;
  for (const ConstPatch& cp : patches) {
    if (cp.fn != cur) {
      cur = cp.fn;
      eval_nodes(efn(cur), p, nullptr, val);
    }
    const uint64_t v = val[cp.node];
    uint64_t w = cp.inverse ? ffhost::div(1 % p, v, p) : v;
    if (cp.negate) w = ffhost::neg(w, p);
    Instr& in = tape[cp.instr];
    in.imm = w;
    // the fast dispatch code of a literal power covers exponents below p/2 only (FFInt::pow's direct branch);
    // a residue above that means a negative power and takes the general code (inversion first)
    if (in.op == OP_POWU) in.fast = (w < (p >> 1) && in.flags == 0) ? FX_POWU : FX_GENERIC;
  }
  for (Collection& pc : pre)
    if (!pc.resolve_constants(p, err)) return false;
  prime = p;
  return true;
}

}  // namespace ffgpu
