"""firefly_b200 -- thin ctypes binding of libffgpu.so (include/ffgpu.h).

The product is the shared library: a tape compiler for FireFly's expression
collections and the sm_100a kernels that evaluate them at batches of probe points
mod a 63-bit prime.  This module only loads it and wraps the C ABI for tests and
bench.py; the C++ drop-in for FireFly's ``BlackBoxBase`` is
``include/firefly_gpu/GpuBlackBox.hpp``.

Mirrors, by name and meaning, the part of the reference API that sits on the hot path
(reference: source/include/firefly/ShuntingYardParser.hpp:44-135):

    ShuntingYardParser(file | functions, vars, check_is_equal)  -> GpuParser(...)
    parser.add_otf / parse_function / add_otf_precompute / trim -> the same names on GpuParser
    parser.precompute_tokens() after FFInt::set_new_prime(p)    -> parser.set_prime(p)
    parser.evaluate_pre(values)                                 -> parser.evaluate_pre(values)
    parser.get_rp_function(i)                                   -> parser.get_rp_function(i)

There is no CPU fallback: importing works anywhere the library was built, evaluating
needs a B200.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional, Sequence

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("FFGPU_LIB") or os.path.join(_HERE, "lib", "libffgpu.so")   # FFGPU_LIB: A/B builds

INSTR_DTYPE = np.dtype(
    [("op", "u1"), ("src", "u1"), ("flags", "u1"), ("fast", "u1"), ("arg", "<u4"), ("imm", "<u8")]
)
# tape.hpp
OPS = ["HALT", "LOAD", "ADD", "SUB", "MUL", "SQR", "NEG", "INV", "POWU", "NPOWU", "POWL", "NPOWL", "STORE", "MULADD",
       "SAVE", "ACCUM", "TERM", "ACCUM2", "ADDVARS", "MULVARS"]
SRCS = ["NONE", "POP", "VAR", "NEGVAR", "CONST", "TMP"]
F_PUSH, F_SWAP, F_NEGR, F_INIT, F_SWAP2, F_INIT2, F_CMUL = 1, 2, 4, 8, 16, 32, 128
TM_LOAD, TM_ADD, TM_SUB, TM_RSUB, TM_MUL = 0, 1, 2, 3, 4
CHECK_IS_EQUAL = 1
NO_CSE = 2
NO_FUSE = 4
NO_CROSS = 16
NO_TABLE = 32
NO_SPLIT = 64
NO_HORNER = 128
ADD_PLAIN, ADD_VALIDATE, ADD_NO_DUPLICATES = 1, 2, 4


class FFGpuError(RuntimeError):
    pass


def _load() -> C.CDLL:
    if not os.path.exists(LIB_PATH):
        raise FFGpuError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(there is no CPU fallback)"
        )
    lib = C.CDLL(LIB_PATH)
    u64p = C.POINTER(C.c_uint64)
    strs = C.POINTER(C.c_char_p)
    sig = {
        "ffgpu_create": (C.c_void_p, [C.POINTER(C.c_int), C.c_int]),
        "ffgpu_destroy": (None, [C.c_void_p]),
        "ffgpu_n_devices": (C.c_int, [C.c_void_p]),
        "ffgpu_last_error": (C.c_char_p, []),
        "ffgpu_compile": (C.c_void_p, [C.c_void_p, strs, C.c_size_t, strs, C.c_size_t, C.c_uint]),
        "ffgpu_compile_text": (C.c_void_p, [C.c_void_p, C.c_char_p, strs, C.c_size_t, C.c_uint]),
        "ffgpu_compile_file": (C.c_void_p, [C.c_void_p, C.c_char_p, strs, C.c_size_t, C.c_uint]),
        "ffgpu_tape_clone": (C.c_void_p, [C.c_void_p]),
        "ffgpu_tape_destroy": (None, [C.c_void_p]),
        "ffgpu_n_vars": (C.c_size_t, [C.c_void_p]),
        "ffgpu_n_functions": (C.c_size_t, [C.c_void_p]),
        "ffgpu_n_out": (C.c_size_t, [C.c_void_p]),
        "ffgpu_n_instr": (C.c_size_t, [C.c_void_p]),
        "ffgpu_stack_depth": (C.c_size_t, [C.c_void_p]),
        "ffgpu_n_temps": (C.c_size_t, [C.c_void_p]),
        "ffgpu_n_slots": (C.c_size_t, [C.c_void_p]),
        "ffgpu_cse_fallbacks": (C.c_size_t, [C.c_void_p]),
        "ffgpu_n_structural_duplicates": (C.c_size_t, [C.c_void_p]),
        "ffgpu_n_shared_denominators": (C.c_size_t, [C.c_void_p]),
        "ffgpu_n_table_rows": (C.c_size_t, [C.c_void_p]),
        "ffgpu_get_table_tape": (C.c_size_t, [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t]),
        "ffgpu_n_table_levels": (C.c_size_t, [C.c_void_p]),
        "ffgpu_n_table_rows_level": (C.c_size_t, [C.c_void_p, C.c_size_t]),
        "ffgpu_table_row_offset": (C.c_size_t, [C.c_void_p, C.c_size_t]),
        "ffgpu_n_split_rows": (C.c_size_t, [C.c_void_p]),
        "ffgpu_rpn": (C.c_size_t, [C.c_void_p, C.c_size_t, C.c_char_p, C.c_size_t]),
        "ffgpu_get_tape": (C.c_size_t, [C.c_void_p, C.c_void_p, C.c_size_t]),
        "ffgpu_work": (None, [C.c_void_p, u64p]),
        "ffgpu_tape_work": (None, [C.c_void_p, u64p]),
        "ffgpu_set_factors": (C.c_int, [C.c_void_p, C.POINTER(C.c_uint32), strs, C.c_size_t, strs, C.c_size_t]),
        "ffgpu_add_function": (C.c_long, [C.c_void_p, C.c_char_p, C.c_uint]),
        "ffgpu_add_function_rpn": (C.c_long, [C.c_void_p, strs, C.c_size_t]),
        "ffgpu_trim": (C.c_int, [C.c_void_p, u64p, C.c_size_t, u64p]),
        "ffgpu_is_stale": (C.c_int, [C.c_void_p]),
        "ffgpu_set_prime": (C.c_int, [C.c_void_p, C.c_uint64]),
        "ffgpu_get_prime": (C.c_uint64, [C.c_void_p]),
        "ffgpu_n_primes": (C.c_uint32, []),
        "ffgpu_prime": (C.c_uint64, [C.c_uint32]),
        "ffgpu_eval": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_uint32]),
        "ffgpu_eval_line": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_uint32]),
        "ffgpu_eval_wire": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_uint32]),
        "ffgpu_eval_device": (
            C.c_int,
            [C.c_void_p, C.c_int, C.c_void_p, C.c_size_t, C.c_size_t, C.c_void_p, C.c_size_t, C.c_void_p],
        ),
        "ffgpu_queue_create": (C.c_void_p, [C.c_void_p, C.c_size_t]),
        "ffgpu_queue_destroy": (None, [C.c_void_p]),
        "ffgpu_queue_push": (C.c_uint64, [C.c_void_p, C.c_void_p]),
        "ffgpu_queue_push_many": (C.c_uint64, [C.c_void_p, C.c_void_p, C.c_size_t]),
        "ffgpu_queue_push_line": (C.c_uint64, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]),
        "ffgpu_queue_pop": (C.c_long, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_int]),
        "ffgpu_queue_pending": (C.c_size_t, [C.c_void_p]),
        "ffgpu_queue_stats": (None, [C.c_void_p, u64p]),
        "ffgpu_queue_error": (C.c_char_p, [C.c_void_p]),
        "ffgpu_host_alloc": (C.c_void_p, [C.c_size_t]),
        "ffgpu_host_free": (None, [C.c_void_p]),
        "ffgpu_set_launch": (C.c_int, [C.c_void_p, C.c_int, C.c_int]),
        "ffgpu_set_chunk": (C.c_int, [C.c_void_p, C.c_size_t]),
        "ffgpu_set_coalesce": (C.c_int, [C.c_void_p, C.c_int, C.c_size_t]),
        "ffgpu_coalesce_stats": (None, [C.c_void_p, C.POINTER(C.c_uint64)]),
        "ffgpu_launch_count": (C.c_uint64, []),
        "ffgpu_fill_points": (
            C.c_int,
            [C.c_void_p, C.c_int, C.c_void_p, C.c_size_t, C.c_size_t, C.c_size_t, C.c_uint64, C.c_uint64, C.c_uint64,
             C.c_uint64, C.c_void_p],
        ),
        "ffgpu_checksum": (
            C.c_int,
            [C.c_void_p, C.c_int, C.c_void_p, C.c_size_t, C.c_size_t, C.c_size_t, C.c_uint64, C.c_void_p, C.c_void_p],
        ),
        "ffgpu_field_ops": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_uint64]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)      # AttributeError here = the library does not match include/ffgpu.h
        fn.restype = res
        fn.argtypes = args
    return lib


lib = _load()
EXPORTS = [
    "ffgpu_create", "ffgpu_destroy", "ffgpu_n_devices", "ffgpu_last_error", "ffgpu_compile", "ffgpu_compile_text",
    "ffgpu_compile_file", "ffgpu_tape_clone", "ffgpu_tape_destroy", "ffgpu_n_vars", "ffgpu_n_functions", "ffgpu_n_out", "ffgpu_n_instr",
    "ffgpu_stack_depth", "ffgpu_n_temps", "ffgpu_n_slots", "ffgpu_cse_fallbacks", "ffgpu_n_structural_duplicates", "ffgpu_n_shared_denominators", "ffgpu_n_table_rows", "ffgpu_get_table_tape", "ffgpu_n_table_levels", "ffgpu_n_table_rows_level", "ffgpu_table_row_offset", "ffgpu_n_split_rows", "ffgpu_rpn", "ffgpu_get_tape", "ffgpu_work", "ffgpu_tape_work", "ffgpu_set_factors", "ffgpu_set_prime", "ffgpu_get_prime",
    "ffgpu_add_function", "ffgpu_add_function_rpn", "ffgpu_trim", "ffgpu_is_stale",
    "ffgpu_n_primes", "ffgpu_prime", "ffgpu_eval", "ffgpu_eval_line", "ffgpu_eval_wire", "ffgpu_eval_device", "ffgpu_host_alloc", "ffgpu_host_free",
    "ffgpu_queue_create", "ffgpu_queue_destroy", "ffgpu_queue_push", "ffgpu_queue_push_many", "ffgpu_queue_push_line",
    "ffgpu_queue_pop", "ffgpu_queue_pending", "ffgpu_queue_stats", "ffgpu_queue_error",
    "ffgpu_set_launch", "ffgpu_set_coalesce", "ffgpu_coalesce_stats", "ffgpu_set_chunk", "ffgpu_launch_count", "ffgpu_fill_points", "ffgpu_checksum",
    "ffgpu_field_ops",
]


def last_error() -> str:
    return (lib.ffgpu_last_error() or b"").decode()


def primes(n: int = 300) -> list[int]:
    """The n largest primes below 2^63, descending (reference: primes(), source/ReconstHelper.cpp:24-125)."""
    return [lib.ffgpu_prime(i) for i in range(min(n, lib.ffgpu_n_primes()))]


def _strs(xs: Sequence[str]):
    arr = (C.c_char_p * max(len(xs), 1))()
    for i, x in enumerate(xs):
        arr[i] = x.encode()
    return arr


class Context:
    """A set of GPUs (default: the current CUDA device)."""

    def __init__(self, devices: Optional[Sequence[int]] = None):
        if devices:
            ids = (C.c_int * len(devices))(*devices)
            self.h = lib.ffgpu_create(ids, len(devices))
        else:
            self.h = lib.ffgpu_create(None, 0)
        if not self.h:
            raise FFGpuError(last_error())

    @property
    def n_devices(self) -> int:
        return lib.ffgpu_n_devices(self.h)

    def close(self):
        if getattr(self, "h", None):
            lib.ffgpu_destroy(self.h)
            self.h = None

    def field_ops(self, op: int, a, b, p: int) -> np.ndarray:
        a = np.ascontiguousarray(a, dtype=np.uint64)
        b = np.ascontiguousarray(b, dtype=np.uint64)
        r = np.empty_like(a)
        if lib.ffgpu_field_ops(self.h, op, a.ctypes.data, b.ctypes.data, r.ctypes.data, a.size, p):
            raise FFGpuError(last_error())
        return r


class GpuParser:
    """A compiled collection (ShuntingYardParser's role on the hot path).

    ``ctx=None`` gives a host-only handle: RPN, tape and work counts can be inspected,
    evaluation raises.
    """

    def __init__(self, ctx: Optional[Context], functions=None, vars: Sequence[str] = (), check_is_equal: bool = False,
                 *, file: Optional[str] = None, text: Optional[str] = None, cse: bool = True,
                 temps: int = 0, fuse: bool = True, cross: bool = True, table: bool = True, split: bool = True, horner: bool = True):
        self.ctx = ctx
        ch = ctx.h if ctx is not None else None
        flags = (CHECK_IS_EQUAL if check_is_equal else 0) | (0 if cse else NO_CSE) | (0 if fuse else NO_FUSE) | (0 if cross else NO_CROSS) | (0 if table else NO_TABLE) | (0 if split else NO_SPLIT) | (0 if horner else NO_HORNER) | ((temps & 15) << 8)
        v = _strs(vars)
        if file is not None:
            self.h = lib.ffgpu_compile_file(ch, file.encode(), v, len(vars), flags)
        elif text is not None:
            self.h = lib.ffgpu_compile_text(ch, text.encode(), v, len(vars), flags)
        else:
            f = _strs(functions)
            self.h = lib.ffgpu_compile(ch, f, len(functions), v, len(vars), flags)
        if not self.h:
            raise FFGpuError(last_error())
        self.n_vars = lib.ffgpu_n_vars(self.h)
        self.n_out = lib.ffgpu_n_out(self.h)
        self.n_functions = lib.ffgpu_n_functions(self.h)

    def close(self):
        if getattr(self, "h", None):
            lib.ffgpu_tape_destroy(self.h)
            self.h = None

    def __del__(self):
        # A host-only handle (no context, no device memory) is released when its last reference goes: loops that
        # compile thousands of collections for inspection need no close().  A handle with a context is released by
        # close() only -- its device buffers must go before the context does, and the order in which the collector
        # visits the two is not ours to choose.
        try:
            if getattr(self, "ctx", None) is None and getattr(self, "h", None) and lib is not None:
                lib.ffgpu_tape_destroy(self.h)
                self.h = None
        except Exception:
            pass

    # -- inspection ------------------------------------------------------------------
    def get_rp_function(self, i: int) -> list[str]:
        n = lib.ffgpu_rpn(self.h, i, None, 0)
        buf = C.create_string_buffer(n + 1)
        lib.ffgpu_rpn(self.h, i, buf, n + 1)
        s = buf.value.decode()
        return s.split(" ") if s else []

    def tape(self) -> np.ndarray:
        n = lib.ffgpu_get_tape(self.h, None, 0)
        arr = np.zeros(n, dtype=INSTR_DTYPE)
        lib.ffgpu_get_tape(self.h, arr.ctypes.data, n)
        return arr

    def table_tape(self, level: int = 0) -> np.ndarray:
        """the tape of the shared rows of one level (its output r = row offset + r = a variable of tape() and of the
        levels after it); empty if there is no such level"""
        n = lib.ffgpu_get_table_tape(self.h, level, None, 0)
        arr = np.zeros(n, dtype=INSTR_DTYPE)
        if n:
            lib.ffgpu_get_table_tape(self.h, level, arr.ctypes.data, n)
        return arr

    def table_levels(self) -> list:
        """[(tape, rows, first row)] of every level of shared rows, in evaluation order"""
        return [(self.table_tape(l), lib.ffgpu_n_table_rows_level(self.h, l), lib.ffgpu_table_row_offset(self.h, l))
                for l in range(lib.ffgpu_n_table_levels(self.h))]

    @property
    def n_split_rows(self) -> int:
        return lib.ffgpu_n_split_rows(self.h)

    @property
    def n_table_rows(self) -> int:
        return lib.ffgpu_n_table_rows(self.h)

    def work(self) -> dict:
        w = (C.c_uint64 * 5)()
        lib.ffgpu_work(self.h, w)
        return {"tokens": w[0], "mm": w[1], "ma": w[2], "inv": w[3], "max_rpn_depth": w[4]}

    def tape_work(self) -> dict:
        """what the compiled tape executes per probe (after folding and sharing)"""
        w = (C.c_uint64 * 5)()
        lib.ffgpu_tape_work(self.h, w)
        return {"instr": w[0], "mm": w[1], "ma": w[2], "inv": w[3], "var_reads": w[4]}

    @property
    def n_instr(self) -> int:
        return lib.ffgpu_n_instr(self.h)

    @property
    def stack_depth(self) -> int:
        return lib.ffgpu_stack_depth(self.h)

    @property
    def n_temps(self) -> int:
        return lib.ffgpu_n_temps(self.h)

    @property
    def cse_fallbacks(self) -> int:
        return lib.ffgpu_cse_fallbacks(self.h)

    @property
    def n_structural_duplicates(self) -> int:
        return lib.ffgpu_n_structural_duplicates(self.h)

    @property
    def n_shared_denominators(self) -> int:
        return lib.ffgpu_n_shared_denominators(self.h)

    @property
    def n_slots(self) -> int:
        return lib.ffgpu_n_slots(self.h)

    def set_factors(self, factors: dict, factor_vars: Sequence[str]):
        """{output index: factor expression}: output i becomes f_i / factor_i on the device (what Reconstructor does
        with parsed_factors after bb.eval, Reconstructor.hpp:3449-3457).  Call set_prime again afterwards."""
        idx = (C.c_uint32 * max(len(factors), 1))(*factors.keys())
        if lib.ffgpu_set_factors(self.h, idx, _strs(list(factors.values())), len(factors), _strs(factor_vars), len(factor_vars)):
            raise FFGpuError(last_error())
        self.n_out = lib.ffgpu_n_out(self.h)
        self.n_functions = lib.ffgpu_n_functions(self.h)

    # -- changing the collection (ShuntingYardParser.cpp:315-377, 720-776) -----------------
    def _refresh(self):
        self.n_out = lib.ffgpu_n_out(self.h)
        self.n_functions = lib.ffgpu_n_functions(self.h)

    def add_otf(self, fun: str, no_duplicates: bool = False) -> int:
        """Append a function; returns its position (its twin's on a check_is_equal parser).  set_prime rebuilds."""
        pos = lib.ffgpu_add_function(self.h, fun.encode(), ADD_NO_DUPLICATES if no_duplicates else 0)
        if pos < 0:
            raise FFGpuError(last_error())
        self._refresh()
        return pos

    def parse_function(self, fun: str, validate_fun: bool = False) -> int:
        pos = lib.ffgpu_add_function(self.h, fun.encode(), ADD_PLAIN | (ADD_VALIDATE if validate_fun else 0))
        if pos < 0:
            raise FFGpuError(last_error())
        self._refresh()
        return pos

    def add_otf_precompute(self, rpn_fun: Sequence[str]) -> int:
        pos = lib.ffgpu_add_function_rpn(self.h, _strs(rpn_fun), len(rpn_fun))
        if pos < 0:
            raise FFGpuError(last_error())
        self._refresh()
        return pos

    def trim(self, elements_to_keep) -> dict:
        """Keep the listed functions (output positions with check_is_equal); returns {old position: new position}."""
        keep = sorted(set(int(k) for k in elements_to_keep))
        arr = (C.c_uint64 * max(len(keep), 1))(*keep)
        new = (C.c_uint64 * max(len(keep), 1))()
        if lib.ffgpu_trim(self.h, arr, len(keep), new):
            raise FFGpuError(last_error())
        self._refresh()
        return {k: new[i] for i, k in enumerate(keep) if new[i] != 2 ** 64 - 1}

    def clone(self) -> "GpuParser":
        """An independent copy: same functions and prime, device buffers of its own (a parser copied by value)."""
        h = lib.ffgpu_tape_clone(self.h)
        if not h:
            raise FFGpuError(last_error())
        q = object.__new__(GpuParser)
        q.ctx, q.h, q.n_vars = self.ctx, h, self.n_vars
        q._refresh()
        return q

    @property
    def stale(self) -> bool:
        return bool(lib.ffgpu_is_stale(self.h))

    # -- prime -----------------------------------------------------------------------
    def set_prime(self, p: int):
        """FFInt::set_new_prime(p) followed by prime_changed() -> precompute_tokens()."""
        if lib.ffgpu_set_prime(self.h, p):
            raise FFGpuError(last_error())

    @property
    def prime(self) -> int:
        return lib.ffgpu_get_prime(self.h)

    # -- evaluation ------------------------------------------------------------------
    def evaluate_pre(self, values, thread_id: int = 0, out: Optional[np.ndarray] = None) -> np.ndarray:
        """values: uint64 [n_vars][n_points] (or [n_vars] for one point) -> uint64 [n_out][n_points]."""
        a = np.asarray(values, dtype=np.uint64)
        single = a.ndim == 1
        if single:
            a = a.reshape(self.n_vars, 1)
        if a.shape[0] != self.n_vars:
            raise ValueError(f"expected {self.n_vars} rows of values, got {a.shape[0]}")
        a = np.ascontiguousarray(a)
        n = a.shape[1]
        if out is None:
            out = np.empty((self.n_out, n), dtype=np.uint64)
        if lib.ffgpu_eval(self.h, a.ctypes.data, n, out.ctypes.data, thread_id):
            raise FFGpuError(last_error())
        return out[:, 0] if single else out

    def evaluate_line(self, a, s, t, thread_id: int = 0, out: Optional[np.ndarray] = None) -> np.ndarray:
        """Probes on the line x_v = a[v]*t + s[v] (Reconstructor::queue_probes, Reconstructor.hpp:2946-2960);
        only t is copied to the GPU.  a, s: [n_vars]; t: [n_points] -> uint64 [n_out][n_points]."""
        a = np.ascontiguousarray(a, dtype=np.uint64)
        s = np.ascontiguousarray(s, dtype=np.uint64)
        t = np.ascontiguousarray(t, dtype=np.uint64)
        if a.shape != (self.n_vars,) or s.shape != (self.n_vars,):
            raise ValueError(f"a and s must have {self.n_vars} entries")
        n = t.shape[0]
        if out is None:
            out = np.empty((self.n_out, n), dtype=np.uint64)
        if lib.ffgpu_eval_line(self.h, a.ctypes.data, s.ctypes.data, t.ctypes.data, n, out.ctypes.data, thread_id):
            raise FFGpuError(last_error())
        return out

    def eval_line_ptr(self, a, s, t_ptr: int, n_points: int, out_ptr: int, thread_id: int = 0):
        a = np.ascontiguousarray(a, dtype=np.uint64)
        s = np.ascontiguousarray(s, dtype=np.uint64)
        if lib.ffgpu_eval_line(self.h, a.ctypes.data, s.ctypes.data, t_ptr, n_points, out_ptr, thread_id):
            raise FFGpuError(last_error())

    def evaluate_wire(self, values_msg, thread_id: int = 0) -> np.ndarray:
        """MPIWorker's message format: [index, x_1..x_n] * k -> [index, f_1..f_m] * k (MPIWorker.hpp:213-228, 479-501)"""
        msg = np.ascontiguousarray(values_msg, dtype=np.uint64).ravel()
        if msg.size % (self.n_vars + 1):
            raise ValueError("Corrupted values: the message length is not a multiple of n_vars + 1")
        k = msg.size // (self.n_vars + 1)
        res = np.empty(k * (self.n_out + 1), dtype=np.uint64)
        if lib.ffgpu_eval_wire(self.h, msg.ctypes.data, k, res.ctypes.data, thread_id):
            raise FFGpuError(last_error())
        return res

    def eval_ptr(self, values_ptr: int, n_points: int, out_ptr: int, thread_id: int = 0):
        """Host pointers (e.g. pinned memory from ``host_alloc``)."""
        if lib.ffgpu_eval(self.h, values_ptr, n_points, out_ptr, thread_id):
            raise FFGpuError(last_error())

    def eval_device(self, d_values: int, in_stride: int, n_points: int, d_out: int, out_stride: int,
                    stream: int = 0, device_index: int = 0):
        """Device pointers; asynchronous on ``stream`` (a raw cudaStream_t)."""
        if lib.ffgpu_eval_device(self.h, device_index, d_values, in_stride, n_points, d_out, out_stride, stream):
            raise FFGpuError(last_error())

    def set_launch(self, probes_per_thread: int = 0, threads_per_block: int = 0):
        if lib.ffgpu_set_launch(self.h, probes_per_thread, threads_per_block):
            raise FFGpuError(last_error())

    def set_chunk(self, points: int = 0):
        lib.ffgpu_set_chunk(self.h, points)

    def set_coalesce(self, enable: int = 1, max_points: int = 0):
        """aggregator of concurrent small calls (include/ffgpu.h, ffgpu_set_coalesce)"""
        if lib.ffgpu_set_coalesce(self.h, enable, max_points):
            raise FFGpuError(last_error())

    def coalesce_stats(self) -> dict:
        s = (C.c_uint64 * 3)()
        lib.ffgpu_coalesce_stats(self.h, s)
        return {"launches": s[0], "calls": s[1], "points": s[2]}


class ProbeQueue:
    """requested_probes / computed_probes around a GpuParser with one GPU-feeding consumer
    (Reconstructor.hpp:245-246, 2962-2971, 2996-3012, 3326-3477; see include/ffgpu.h, probe queue)."""

    def __init__(self, parser: GpuParser, max_batch: int = 0):
        self.parser = parser
        self.h = lib.ffgpu_queue_create(parser.h, max_batch)
        if not self.h:
            raise FFGpuError("could not create the probe queue")

    def close(self):
        if getattr(self, "h", None):
            lib.ffgpu_queue_destroy(self.h)
            self.h = None

    def push(self, values) -> int:
        v = np.ascontiguousarray(values, dtype=np.uint64)
        if v.shape != (self.parser.n_vars,):
            raise ValueError(f"expected {self.parser.n_vars} values")
        return lib.ffgpu_queue_push(self.h, v.ctypes.data)

    def push_many(self, points) -> int:
        """points: [n][n_vars] (point-major, like the reference's queue entries); returns the first index"""
        v = np.ascontiguousarray(points, dtype=np.uint64)
        if v.ndim != 2 or v.shape[1] != self.parser.n_vars:
            raise ValueError(f"expected [n][{self.parser.n_vars}] values")
        return lib.ffgpu_queue_push_many(self.h, v.ctypes.data, v.shape[0])

    def push_line(self, a, s, t) -> int:
        a = np.ascontiguousarray(a, dtype=np.uint64)
        s = np.ascontiguousarray(s, dtype=np.uint64)
        t = np.ascontiguousarray(t, dtype=np.uint64)
        if a.shape != (self.parser.n_vars,) or s.shape != (self.parser.n_vars,):
            raise ValueError(f"a and s must have {self.parser.n_vars} entries")
        return lib.ffgpu_queue_push_line(self.h, a.ctypes.data, s.ctypes.data, t.ctypes.data, t.shape[0])

    def pop(self, cap: int = 4096, wait: bool = True):
        """-> (indices[k], values[n_out][k]) of up to cap finished probes (k = 0: nothing outstanding)"""
        idx = np.empty(cap, dtype=np.uint64)
        out = np.empty((self.parser.n_out, cap), dtype=np.uint64)
        k = lib.ffgpu_queue_pop(self.h, idx.ctypes.data, out.ctypes.data, cap, 1 if wait else 0)
        if k < 0:
            raise FFGpuError((lib.ffgpu_queue_error(self.h) or b"").decode())
        return idx[:k].copy(), out[:, :k].copy()

    @property
    def pending(self) -> int:
        return lib.ffgpu_queue_pending(self.h)

    def stats(self) -> dict:
        s = (C.c_uint64 * 3)()
        lib.ffgpu_queue_stats(self.h, s)
        return {"batches": s[0], "probes": s[1], "largest_batch": s[2]}


def launch_count() -> int:
    return lib.ffgpu_launch_count()


from .points import synthetic_points  # noqa: E402,F401  (pure numpy; kept importable on its own)
