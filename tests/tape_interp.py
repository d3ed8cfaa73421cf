"""Host interpretation of the DEVICE tape with Python integers -- TEST INFRASTRUCTURE.

Lets the CPU test-suite check the tape compiler (firefly_b200/csrc/compiler.cpp) against the
oracle without a GPU: same instruction semantics as the sm_100a interpreter
(firefly_b200/csrc/kernels.cu), one point at a time, exact arithmetic.
Field semantics: reference FFInt DEFAULT arithmetic (source/FFInt.cpp:130-192,226-228,356-405).
"""
import numpy as np

from firefly_b200 import (F_CMUL, F_INIT, F_INIT2, F_NEGR, F_PUSH, F_SWAP, F_SWAP2, OPS, SRCS, TM_ADD, TM_LOAD, TM_MUL,
                          TM_RSUB, TM_SUB)

OP = {n: i for i, n in enumerate(OPS)}
SRC = {n: i for i, n in enumerate(SRCS)}


def inv(a, p):
    return pow(a, p - 2, p) if a % p else 0


def pow_ff(b, e, p):
    """FFInt::pow(const FFInt&)"""
    if e == 2:
        return b * b % p
    if e < (p >> 1):
        return pow(b, e, p)
    return pow(inv(b, p), p - e - 1, p)


def to_neg_int(n):
    v = (-(n & 0xFFFFFFFF)) & 0xFFFFFFFF
    return v - (1 << 32) if v >= (1 << 31) else v


def pow_int(b, e, p):
    """FFInt::pow<int>"""
    if e == 2:
        return b * b % p
    if e >= 0:
        return pow(b, e, p)
    return pow(inv(b, p), -e, p)


def run(tape: np.ndarray, point, p: int, n_out: int):
    """Evaluate the whole tape at one point (sequence of ints); returns the list of outputs."""
    out = [None] * n_out
    acc = 0
    stack = []
    temps = {}
    for ins in tape:
        op, src, flags, arg, imm = int(ins["op"]), int(ins["src"]), int(ins["flags"]), int(ins["arg"]), int(ins["imm"])
        if op == OP["HALT"]:
            break
        b = None
        if op >= OP["TERM"]:
            # fused forms decode their own operands
            if op == OP["TERM"]:
                mode = (flags >> 4) & 7
                t = pow_ff(point[arg & 0xFFFF] % p, arg >> 16, p)
                if flags & F_CMUL:
                    t = t * imm % p
                if mode == TM_LOAD:
                    if flags & F_PUSH:
                        stack.append(acc)
                    acc = t
                elif mode == TM_ADD:
                    acc = (acc + t) % p
                elif mode == TM_SUB:
                    acc = (acc - t) % p
                elif mode == TM_RSUB:
                    acc = (t - acc) % p
                elif mode == TM_MUL:
                    acc = acc * t % p
                else:
                    raise ValueError("bad TERM mode")
            elif op == OP["ACCUM2"]:
                for slot, sw, init in ((arg & 0xFF, flags & F_SWAP, flags & F_INIT), ((arg >> 8) & 0xFF, flags & F_SWAP2, flags & F_INIT2)):
                    v = (-acc) % p if sw else acc
                    temps[slot] = v if init else (temps[slot] + v) % p
            elif op in (OP["ADDVARS"], OP["MULVARS"]):
                for k in range(arg & 0xFF):
                    v = point[(imm >> (8 * k)) & 0xFF] % p
                    if op == OP["MULVARS"]:
                        acc = acc * v % p
                    elif (arg >> (8 + k)) & 1:
                        acc = (acc - v) % p
                    else:
                        acc = (acc + v) % p
            else:
                raise ValueError(f"bad opcode {op}")
            continue
        if src == SRC["POP"]:
            b = stack.pop()
        elif src == SRC["VAR"]:
            b = point[arg] % p
        elif src == SRC["NEGVAR"]:
            b = (-point[arg]) % p
        elif src == SRC["CONST"]:
            b = imm
        elif src == SRC["TMP"]:
            b = temps[arg]
        if op == OP["LOAD"]:
            if flags & F_PUSH:
                stack.append(acc)
            acc = b
        elif op == OP["ADD"]:
            acc = (acc + b) % p
        elif op == OP["SUB"]:
            acc = (b - acc) % p if flags & F_SWAP else (acc - b) % p
        elif op == OP["MUL"]:
            acc = acc * b % p
        elif op == OP["SQR"]:
            acc = acc * acc % p
        elif op == OP["NEG"]:
            acc = (-acc) % p
        elif op == OP["INV"]:
            acc = inv(acc, p)
        elif op == OP["POWU"]:
            acc = pow_ff(acc, imm, p)
        elif op == OP["NPOWU"]:
            acc = pow_int(acc, to_neg_int(imm), p)
        elif op == OP["POWL"]:
            acc = pow_ff(b, acc, p) if flags & F_SWAP else pow_ff(acc, b, p)
        elif op == OP["NPOWL"]:
            acc = pow_int(b, to_neg_int(acc), p) if flags & F_SWAP else pow_int(acc, to_neg_int(b), p)
        elif op == OP["STORE"]:
            out[arg] = acc
        elif op == OP["MULADD"]:
            x = point[arg] % p
            acc = (imm - acc * x) % p if flags & F_SWAP else (acc * x + imm) % p
        elif op == OP["SAVE"]:
            temps[arg] = acc
        elif op == OP["ACCUM"]:
            v = (-acc) % p if flags & F_SWAP else acc
            temps[arg] = v if flags & F_INIT else (temps[arg] + v) % p
        else:
            raise ValueError(f"bad opcode {op}")
        if flags & F_NEGR:
            acc = (-acc) % p
    assert not stack, "spill stack not empty at HALT"
    return out
