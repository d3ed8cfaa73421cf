"""ctypes access to the checkers -- TEST INFRASTRUCTURE.

  * oracle/libfforacle.so   the C restatement of the reference's hot path (oracle/ff_oracle.c)
  * oracle/_ref/libfirefly_ref.so   the UNMODIFIED reference compiled by oracle/Makefile (`make ref`);
    exists only where /root/reference was available at build time (it travels to the GPU box as a
    built file).
"""
import ctypes as C
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_SO = os.path.join(ROOT, "oracle", "_ref", "libfirefly_ref.so")
ORACLE_SO = os.path.join(ROOT, "oracle", "libfforacle.so")


def _strs(xs):
    arr = (C.c_char_p * max(len(xs), 1))()
    for i, x in enumerate(xs):
        arr[i] = x.encode()
    return arr


class Ref:
    """The reference's own ShuntingYardParser behind oracle/ref_wrapper.cpp.  FFInt::p is a process global."""

    def __init__(self):
        self.lib = C.CDLL(REF_SO)
        L = self.lib
        L.ffref_set_prime.argtypes = [C.c_uint64]
        L.ffref_get_prime.restype = C.c_uint64
        L.ffref_prime.restype = C.c_uint64
        L.ffref_prime.argtypes = [C.c_uint32]
        L.ffref_n_primes.restype = C.c_uint32
        L.ffref_parse.restype = C.c_void_p
        L.ffref_parse.argtypes = [C.POINTER(C.c_char_p), C.c_size_t, C.POINTER(C.c_char_p), C.c_size_t, C.c_int]
        L.ffref_parse_file.restype = C.c_void_p
        L.ffref_parse_file.argtypes = [C.c_char_p, C.POINTER(C.c_char_p), C.c_size_t, C.c_int]
        L.ffref_free.argtypes = [C.c_void_p]
        L.ffref_n_out.restype = C.c_size_t
        L.ffref_n_out.argtypes = [C.c_void_p]
        L.ffref_n_functions.restype = C.c_size_t
        L.ffref_n_functions.argtypes = [C.c_void_p]
        L.ffref_prime_changed.argtypes = [C.c_void_p]
        L.ffref_rpn.restype = C.c_size_t
        L.ffref_rpn.argtypes = [C.c_void_p, C.c_size_t, C.c_char_p, C.c_size_t]
        L.ffref_eval_scalar.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]
        L.ffref_eval.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_int, C.c_int]
        for name in ("add", "sub", "mul", "div", "pow"):
            f = getattr(L, "ffref_" + name)
            f.restype = C.c_uint64
            f.argtypes = [C.c_uint64, C.c_uint64]
        L.ffref_neg.restype = C.c_uint64
        L.ffref_neg.argtypes = [C.c_uint64]
        L.ffref_pow_int.restype = C.c_uint64
        L.ffref_pow_int.argtypes = [C.c_uint64, C.c_int]
        L.ffref_to_neg_int.restype = C.c_int
        L.ffref_to_neg_int.argtypes = [C.c_uint64]
        strs = C.POINTER(C.c_char_p)
        L.ffref_add_otf.restype = C.c_size_t
        L.ffref_add_otf.argtypes = [C.c_void_p, C.c_char_p, C.c_int]
        L.ffref_parse_function.argtypes = [C.c_void_p, C.c_char_p, strs, C.c_size_t, C.c_int]
        L.ffref_add_otf_precompute.restype = C.c_size_t
        L.ffref_add_otf_precompute.argtypes = [C.c_void_p, strs, C.c_size_t]
        L.ffref_trim.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]
        L.ffref_refresh.argtypes = [C.c_void_p, C.c_int]
        L.ffref_compute_bunch_size.restype = C.c_uint32
        L.ffref_compute_bunch_size.argtypes = [C.c_uint32] * 3
        L.ffref_compute_job_number.restype = C.c_uint32
        L.ffref_compute_job_number.argtypes = [C.c_uint32] * 4

    def set_prime(self, p):
        self.lib.ffref_set_prime(p)

    def parse(self, funs, vars, check_is_equal=False):
        return RefParser(self, self.lib.ffref_parse(_strs(funs), len(funs), _strs(vars), len(vars), int(check_is_equal)),
                         len(vars))

    def parse_file(self, path, vars, check_is_equal=False):
        return RefParser(self, self.lib.ffref_parse_file(path.encode(), _strs(vars), len(vars), int(check_is_equal)),
                         len(vars))


class RefParser:
    def __init__(self, ref, h, n_vars):
        self.ref, self.h, self.n_vars = ref, h, n_vars
        self.n_out = ref.lib.ffref_n_out(h)
        self.n_functions = ref.lib.ffref_n_functions(h)

    def prime_changed(self):
        self.ref.lib.ffref_prime_changed(self.h)

    def rpn(self, i):
        n = self.ref.lib.ffref_rpn(self.h, i, None, 0)
        buf = C.create_string_buffer(n + 1)
        self.ref.lib.ffref_rpn(self.h, i, buf, n + 1)
        s = buf.value.decode()
        return s.split(" ") if s else []

    def eval(self, values, bunch=1, threads=1):
        a = np.ascontiguousarray(values, dtype=np.uint64)
        n = a.shape[1]
        out = np.empty((self.n_out, n), dtype=np.uint64)
        if bunch == 1 and threads == 1:
            self.ref.lib.ffref_eval_scalar(self.h, a.ctypes.data, n, out.ctypes.data)
        else:
            self.ref.lib.ffref_eval(self.h, a.ctypes.data, n, out.ctypes.data, bunch, threads)
        return out

    # the rest of the parser's public interface (ShuntingYardParser.cpp:315-377, 720-776)
    def add_otf(self, fun, no_duplicates=False):
        return self.ref.lib.ffref_add_otf(self.h, fun.encode(), int(no_duplicates))

    def parse_function(self, fun, vars, validate=False):
        self.ref.lib.ffref_parse_function(self.h, fun.encode(), _strs(vars), len(vars), int(validate))

    def add_otf_precompute(self, tokens):
        return self.ref.lib.ffref_add_otf_precompute(self.h, _strs(tokens), len(tokens))

    def trim(self, keep):
        keep = sorted(set(keep))
        a = np.array(keep, dtype=np.uint64)
        new = np.zeros(max(len(keep), 1), dtype=np.uint64)
        self.ref.lib.ffref_trim(self.h, a.ctypes.data, len(keep), new.ctypes.data)
        return {k: int(new[i]) for i, k in enumerate(keep) if int(new[i]) != 2 ** 64 - 1}

    def refresh(self, force=True):
        self.ref.lib.ffref_refresh(self.h, int(force))
        self.n_out = self.ref.lib.ffref_n_out(self.h)
        self.n_functions = self.ref.lib.ffref_n_functions(self.h)

    def close(self):
        if self.h:
            self.ref.lib.ffref_free(self.h)
            self.h = None


def have_ref():
    return os.path.exists(REF_SO)
