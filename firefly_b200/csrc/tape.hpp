// tape.hpp -- the device instruction tape: format shared by the host-side tape
// compiler (compiler.cpp) and the sm_100a interpreter kernels (kernels.cu).
//
// What it replaces.  The reference pre-compiles every function of a collection
// into `precomp_tokens`, a vector of (opcode, FFInt) pairs in reverse Polish order
// (reference: source/include/firefly/ShuntingYardParser.hpp:148,220-238;
// builder source/ShuntingYardParser.cpp:593-718) and runs a std::stack machine over
// it per function and per point (ShuntingYardParser.hpp:365-479).
//
// This tape is NOT that token list.  It is straight-line code for an
// ACCUMULATOR machine generated from the expression tree of each function:
//
//   * every lane (= one probe point) owns an accumulator ACC that lives in
//     registers and a small spill stack that lives in shared memory;
//   * an instruction is `ACC = ACC op SRC` where SRC is a variable of the probe
//     point, a negated variable, a per-prime constant carried inline in the
//     instruction, or the top of the spill stack;
//   * the spill stack is only touched when both operands of a binary operator are
//     compound expressions; the compiler orders sub-trees (Sethi-Ullman) so that
//     the depth stays logarithmic;
//   * values used more than once (equal sub-expressions of one function) are parked
//     in temp slots, also in shared memory, and read back non-destructively;
//   * frequent runs of instructions have fused forms (TERM, ACCUM2, ADDVARS, MULVARS,
//     MULADD) that cost one dispatch of the interpreter;
//   * all functions of a collection are concatenated into one tape; each ends in a
//     STORE that writes the canonical residue to out[slot][point].
//
// The tape's structure is prime independent; only the 64-bit `imm` fields of
// constant-carrying instructions change with the prime (the reference keeps
// `partial_rpn` for the same purpose, ShuntingYardParser.cpp:584-590).
#pragma once
#include <cstdint>

namespace ffgpu {

// One instruction = 16 bytes, fetched with a single 128-bit load.
struct alignas(16) Instr {
  uint8_t op;      // Op
  uint8_t src;     // Src
  uint8_t flags;   // F_*
  uint8_t fast;    // Fast: dense dispatch code of the common (op, src, flags) combinations, 0 = generic path
  uint32_t arg;    // variable index (SRC_VAR/SRC_NEGVAR), temp slot (SRC_TMP, SAVE, ACCUM), output slot (STORE);
                   // the fused forms pack their own fields in here (see Op)
  uint64_t imm;    // constant (SRC_CONST, canonical residue mod the current prime)
};
static_assert(sizeof(Instr) == 16, "Instr must be 16 bytes");

enum Op : uint8_t {
  OP_HALT = 0,
  OP_LOAD,    // ACC = SRC                      (F_PUSH: spill the old ACC first)
  OP_ADD,     // ACC = ACC + SRC
  OP_SUB,     // ACC = ACC - SRC                (F_SWAP: ACC = SRC - ACC)
  OP_MUL,     // ACC = ACC * SRC
  OP_SQR,     // ACC = ACC * ACC
  OP_NEG,     // ACC = -ACC
  OP_INV,     // ACC = 1/ACC with 1/0 := 0      (block-batched inversion)
  OP_POWU,    // ACC = ACC ^ e, e = imm a residue, same for all lanes; FFInt::pow(const FFInt&) semantics
  OP_NPOWU,   // ACC = ACC ^ -(int)e, e = imm;   FFInt::pow<int>(to_neg_int()) semantics
  OP_POWL,    // ACC = ACC ^ SRC, exponent differs per lane (F_SWAP: ACC = SRC ^ ACC)
  OP_NPOWL,   // ACC = ACC ^ -(int)SRC per lane  (F_SWAP likewise)
  OP_STORE,   // out[arg][point] = canonical(ACC)
  OP_MULADD,  // ACC = ACC * var[arg] + imm      (Horner step; F_SWAP: ACC = imm - ACC * var[arg])
  OP_SAVE,    // temp[arg] = ACC                 (ACC unchanged)
  OP_ACCUM,   // temp[arg] += ACC  (F_SWAP: -=;  F_INIT: temp[arg] = +-ACC);  ACC unchanged
  // ---- fused forms (compiler.cpp, fuse()): one dispatch for a frequent run of instructions ----------------
  OP_TERM,    // T = var[arg & 0xffff] ^ (arg >> 16)  [* imm if F_CMUL];  then by mode (flags bits 4..6):
              //   TM_LOAD  ACC = T (F_PUSH: spill the old ACC first)      <- LOAD.VAR[.P], POWU, [MUL.CONST]
              //   TM_ADD   ACC = ACC + T   TM_SUB  ACC = ACC - T          <- LOAD.VAR.P, POWU, [MUL.CONST], ADD/SUB.POP
              //   TM_RSUB  ACC = T - ACC   TM_MUL  ACC = ACC * T
  OP_ACCUM2,  // two OP_ACCUMs of the same ACC: slot A = arg & 0xff (F_SWAP, F_INIT), slot B = arg >> 8 (F_SWAP2, F_INIT2)
  OP_ADDVARS, // ACC = ACC +- var[id_0] +- ... : n = arg & 0xff (2..8), bit i of arg >> 8 set = subtract, ids = bytes of imm
  OP_MULVARS, // ACC = ACC * var[id_0] * ... : n = arg & 0xff (2..8), ids = bytes of imm
  OP_COUNT
};

// Dense dispatch codes.  Purely an encoding of (op, src, flags) -- the interpreter's first switch
// goes over these; anything without a code (FX_GENERIC) takes the general two-level decode.
enum Fast : uint8_t {
  FX_GENERIC = 0,
  FX_LOAD_VAR, FX_LOAD_CONST, FX_PUSH_LOAD_VAR, FX_PUSH_LOAD_CONST,
  FX_ADD_VAR, FX_ADD_CONST, FX_ADD_POP,
  FX_SUB_VAR, FX_RSUB_VAR, FX_RSUB_CONST, FX_SUB_POP, FX_RSUB_POP,
  FX_MUL_VAR, FX_MUL_CONST, FX_MUL_POP,
  FX_POWU, FX_INV, FX_STORE, FX_SQR,
  FX_SAVE, FX_ACCUM_INIT, FX_ACCUM_ADD, FX_ACCUM_SUB,
  FX_LOAD_TMP, FX_PUSH_LOAD_TMP, FX_ADD_TMP, FX_MUL_TMP,
  FX_TERM, FX_ACCUM2, FX_ADDVARS, FX_MULVARS, FX_MULADD,
  FX_COUNT
};

enum Src : uint8_t {
  SRC_NONE = 0,
  SRC_POP,      // pop the spill stack
  SRC_VAR,      // values[arg][point]
  SRC_NEGVAR,   // -values[arg][point]
  SRC_CONST,    // imm
  SRC_TMP,      // temp slot arg (non-destructive read)
};

enum Flags : uint8_t {
  F_PUSH = 1,   // OP_LOAD: push ACC before loading
  F_SWAP = 2,   // non-commutative ops: exchange the roles of ACC and SRC
  F_NEGR = 4,   // negate the result (the parser's `!` and `;` operators)
  F_INIT = 8,   // OP_ACCUM: first contribution, the slot is overwritten
  F_SWAP2 = 16, // OP_ACCUM2: F_SWAP of the second slot
  F_INIT2 = 32, // OP_ACCUM2: F_INIT of the second slot
  F_CMUL = 128, // OP_TERM: multiply the power by imm
};
// OP_TERM: what happens with the term, in flags bits 4..6
enum TermMode : uint8_t { TM_LOAD = 0, TM_ADD = 1, TM_SUB = 2, TM_RSUB = 3, TM_MUL = 4 };
constexpr uint8_t TM_SHIFT = 4;

}  // namespace ffgpu
