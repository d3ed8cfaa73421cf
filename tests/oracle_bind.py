"""ctypes binding of oracle/libfforacle.so (oracle/ff_oracle.h) -- TEST INFRASTRUCTURE."""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
ORACLE_SO = os.path.join(ORACLE_DIR, "libfforacle.so")


def _strs(xs):
    arr = (C.c_char_p * max(len(xs), 1))()
    for i, x in enumerate(xs):
        arr[i] = x.encode()
    return arr


class Oracle:
    def __init__(self):
        if not os.path.exists(ORACLE_SO):
            subprocess.run(["make", "-s", "-C", ORACLE_DIR, "all"], check=True)
        L = self.lib = C.CDLL(ORACLE_SO)
        u64 = C.c_uint64
        for name, res, args in [
            ("ffo_add", u64, [u64, u64, u64]), ("ffo_sub", u64, [u64, u64, u64]), ("ffo_neg", u64, [u64, u64]),
            ("ffo_mul", u64, [u64, u64, u64]), ("ffo_inv", u64, [u64, u64]), ("ffo_div", u64, [u64, u64, u64]),
            ("ffo_pow", u64, [u64, u64, u64]), ("ffo_pow_int", u64, [u64, C.c_int32, u64]),
            ("ffo_to_neg_int", C.c_int32, [u64]),
            ("ffo_parse", C.c_void_p, [C.POINTER(C.c_char_p), C.c_size_t, C.POINTER(C.c_char_p), C.c_size_t, C.c_int, u64]),
            ("ffo_parse_text", C.c_void_p, [C.c_char_p, C.POINTER(C.c_char_p), C.c_size_t, C.c_int, u64]),
            ("ffo_free", None, [C.c_void_p]), ("ffo_last_error", C.c_char_p, []),
            ("ffo_set_prime", C.c_int, [C.c_void_p, u64]),
            ("ffo_n_functions", C.c_size_t, [C.c_void_p]), ("ffo_n_out", C.c_size_t, [C.c_void_p]),
            ("ffo_n_vars", C.c_size_t, [C.c_void_p]),
            ("ffo_rpn", C.c_size_t, [C.c_void_p, C.c_size_t, C.c_char_p, C.c_size_t]),
            ("ffo_n_tokens", C.c_size_t, [C.c_void_p, C.c_size_t]),
            ("ffo_token_op", C.c_int, [C.c_void_p, C.c_size_t, C.c_size_t]),
            ("ffo_token_val", u64, [C.c_void_p, C.c_size_t, C.c_size_t]),
            ("ffo_eval_position", u64, [C.c_void_p, C.c_size_t]),
            ("ffo_eval", C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
            ("ffo_compute_bunch_size", C.c_uint32, [C.c_uint32] * 3),
            ("ffo_compute_job_number", C.c_uint32, [C.c_uint32] * 4),
            ("ffo_n_primes", C.c_uint32, []), ("ffo_prime", u64, [C.c_uint32]),
        ]:
            f = getattr(L, name)
            f.restype = res
            f.argtypes = args

    def error(self):
        return (self.lib.ffo_last_error() or b"").decode()

    def parse(self, funs, vars, p, check_is_equal=False):
        h = self.lib.ffo_parse(_strs(funs), len(funs), _strs(vars), len(vars), int(check_is_equal), p)
        if not h:
            raise ValueError(self.error())
        return OracleParser(self, h)

    def parse_text(self, text, vars, p, check_is_equal=False):
        h = self.lib.ffo_parse_text(text.encode(), _strs(vars), len(vars), int(check_is_equal), p)
        if not h:
            raise ValueError(self.error())
        return OracleParser(self, h)


class OracleParser:
    def __init__(self, o, h):
        self.o, self.h = o, h
        L = o.lib
        self.n_functions = L.ffo_n_functions(h)
        self.n_out = L.ffo_n_out(h)
        self.n_vars = L.ffo_n_vars(h)

    def set_prime(self, p):
        if self.o.lib.ffo_set_prime(self.h, p):
            raise ValueError(self.o.error())

    def rpn(self, i):
        n = self.o.lib.ffo_rpn(self.h, i, None, 0)
        buf = C.create_string_buffer(n + 1)
        self.o.lib.ffo_rpn(self.h, i, buf, n + 1)
        s = buf.value.decode()
        return s.split(" ") if s else []

    def tokens(self, i):
        L = self.o.lib
        return [(L.ffo_token_op(self.h, i, j), L.ffo_token_val(self.h, i, j)) for j in range(L.ffo_n_tokens(self.h, i))]

    def eval(self, values):
        a = np.ascontiguousarray(values, dtype=np.uint64)
        if a.ndim == 1:
            a = a.reshape(self.n_vars, 1)
        out = np.empty((self.n_out, a.shape[1]), dtype=np.uint64)
        if self.o.lib.ffo_eval(self.h, a.ctypes.data, a.shape[1], out.ctypes.data):
            raise ValueError(self.o.error())
        return out

    def close(self):
        if self.h:
            self.o.lib.ffo_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
