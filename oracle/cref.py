"""ctypes loader for the C++ oracle / CPU-baseline port (oracle/csrc/cref.cpp).

ORACLE -- test infrastructure only (see oracle/__init__.py).  All arrays are numpy uint64 with the
reference's in-memory layout: little-endian u64 limbs, Montgomery form for field elements / points,
canonical form for scalars (FrRepr).
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

OK, UNEXPECTED_IDENTITY, UNEXPECTED_EOF, DEGREE_TOO_LARGE, BAD_ARG = 0, 1, 2, 3, 4


def build(force=False):
    so = os.path.join(_HERE, "libcref.so")
    src = os.path.join(_HERE, "csrc", "cref.cpp")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B", "libcref.so"], stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = ctypes.CDLL(build())
        _LIB.cref_hardware_threads.restype = ctypes.c_int
    return _LIB


def _p(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def _u64(a):
    return np.ascontiguousarray(a, dtype=np.uint64)


def hardware_threads():
    return lib().cref_hardware_threads()


_OPS = dict(add=0, sub=1, mul=2, square=3, double=4, negate=5, into_repr=6, from_repr=7, inverse=8)


def field_vec(field, op, a, b=None):
    a = _u64(a)
    out = np.empty_like(a)
    n = a.shape[0]
    if b is not None:
        b = _u64(b)
    fn = lib().cref_fr_vec if field == "fr" else lib().cref_fq_vec
    fn(ctypes.c_int(_OPS[op]), _p(a), _p(b), _p(out), ctypes.c_size_t(n))
    return out


def scalar_muls(group, base_xy, scalars, threads=0):
    """[k_i] * base -> (xy array (n, 12|24), inf flags (n,))"""
    scalars = _u64(scalars)
    n = scalars.shape[0]
    w = 12 if group == "g1" else 24
    out = np.empty((n, w), dtype=np.uint64)
    inf = np.zeros(n, dtype=np.uint8)
    fn = lib().cref_g1_scalar_muls if group == "g1" else lib().cref_g2_scalar_muls
    fn(_p(_u64(base_xy)), _p(scalars), ctypes.c_size_t(n), _p(out), _p(inf), ctypes.c_int(threads))
    return out, inf


def multiexp(group, bases_xy, scalars, density=None, base_offset=0, inf=None, threads=0):
    """bellman multiexp on the CPU.  Returns (status, jacobian limbs (18|36,))."""
    bases_xy = _u64(bases_xy)
    scalars = _u64(scalars)
    w = 18 if group == "g1" else 36
    out = np.zeros(w, dtype=np.uint64)
    if density is not None:
        density = np.ascontiguousarray(density, dtype=np.uint8)
        assert density.shape[0] == scalars.shape[0]
    if inf is not None:
        inf = np.ascontiguousarray(inf, dtype=np.uint8)
    fn = lib().cref_g1_multiexp if group == "g1" else lib().cref_g2_multiexp
    fn.restype = ctypes.c_int
    st = fn(_p(bases_xy), _p(inf), ctypes.c_size_t(bases_xy.shape[0]), ctypes.c_size_t(base_offset), _p(scalars),
            ctypes.c_size_t(scalars.shape[0]), _p(density), _p(out), ctypes.c_int(threads))
    return st, out


def into_affine(group, jac):
    jac = _u64(jac)
    w = 12 if group == "g1" else 24
    out = np.zeros(w, dtype=np.uint64)
    fn = lib().cref_g1_into_affine if group == "g1" else lib().cref_g2_into_affine
    fn.restype = ctypes.c_int
    inf = fn(_p(jac), _p(out))
    return out, bool(inf)


def point_op(group, op, a_jac, b=None, b_inf=False):
    w = 18 if group == "g1" else 36
    out = np.zeros(w, dtype=np.uint64)
    fn = lib().cref_g1_point_op if group == "g1" else lib().cref_g2_point_op
    code = dict(double=0, add=1, add_mixed=2)[op]
    fn(ctypes.c_int(code), _p(_u64(a_jac)), _p(None if b is None else _u64(b)), ctypes.c_int(int(b_inf)), _p(out))
    return out


FFT, IFFT, COSET_FFT, ICOSET_FFT = 0, 1, 2, 3


def fft(coeffs, kind=FFT, threads=0, serial=False):
    """EvaluationDomain::{fft,ifft,coset_fft,icoset_fft} in place on a copy; coeffs (m,4) Montgomery."""
    a = _u64(coeffs).copy()
    m = a.shape[0]
    log_m = m.bit_length() - 1
    assert 1 << log_m == m
    fn = lib().cref_fft
    fn.restype = ctypes.c_int
    st = fn(_p(a), ctypes.c_uint32(log_m), ctypes.c_int(kind), ctypes.c_int(threads), ctypes.c_int(int(serial)))
    assert st == OK
    return a


def parallel_fft(coeffs, log_cpus, threads=0):
    a = _u64(coeffs).copy()
    log_m = a.shape[0].bit_length() - 1
    fn = lib().cref_parallel_fft
    fn.restype = ctypes.c_int
    st = fn(_p(a), ctypes.c_uint32(log_m), ctypes.c_uint32(log_cpus), ctypes.c_int(threads))
    assert st == OK
    return a


def h_poly(a, b, c, threads=0):
    a, b, c = _u64(a).copy(), _u64(b).copy(), _u64(c).copy()
    m = a.shape[0]
    log_m = m.bit_length() - 1
    out = np.zeros((m - 1, 4), dtype=np.uint64)
    fn = lib().cref_h_poly
    fn.restype = ctypes.c_int
    st = fn(_p(a), _p(b), _p(c), ctypes.c_uint32(log_m), _p(out), ctypes.c_int(threads))
    assert st == OK
    return out


class ResidentBases:
    """A base vector loaded once into the oracle's memory (the `Arc<Vec<G::Affine>>` a bellman caller keeps), so that a timed
    multiexp does not include the conversion of the numpy array."""

    def __init__(self, group, handle, n):
        self.group, self.handle, self.n = group, handle, n

    @classmethod
    def load(cls, group, bases_xy, inf=None):
        bases_xy = _u64(bases_xy)
        fn = lib().cref_g1_bases_load if group == "g1" else lib().cref_g2_bases_load
        fn.restype = ctypes.c_void_p
        if inf is not None:
            inf = np.ascontiguousarray(inf, dtype=np.uint8)
        h = fn(_p(bases_xy), _p(inf), ctypes.c_size_t(bases_xy.shape[0]))
        return cls(group, ctypes.c_void_p(h), bases_xy.shape[0])

    @classmethod
    def walk_g1(cls, start_xy, step_xy, n, n_first=0, threads=0):
        """P_i = start + i * step, i < n, generated on the host cores (synthetic bases for the CPU arm of the bench).
        Returns (handle, first n_first points as (n_first, 12) limbs)."""
        fn = lib().cref_g1_bases_walk
        fn.restype = ctypes.c_void_p
        first = np.zeros((max(n_first, 1), 12), dtype=np.uint64)
        h = fn(_p(_u64(start_xy)), _p(_u64(step_xy)), ctypes.c_size_t(n), _p(first), ctypes.c_size_t(n_first), ctypes.c_int(threads))
        return cls("g1", ctypes.c_void_p(h), n), first[:n_first]

    def multiexp(self, scalars, density=None, base_offset=0, threads=0):
        scalars = _u64(scalars)
        out = np.zeros(18 if self.group == "g1" else 36, dtype=np.uint64)
        if density is not None:
            density = np.ascontiguousarray(density, dtype=np.uint8)
            assert density.shape[0] == scalars.shape[0]
        fn = lib().cref_g1_multiexp_h if self.group == "g1" else lib().cref_g2_multiexp_h
        fn.restype = ctypes.c_int
        st = fn(self.handle, ctypes.c_size_t(base_offset), _p(scalars), ctypes.c_size_t(scalars.shape[0]), _p(density), _p(out), ctypes.c_int(threads))
        return st, out

    def free(self):
        if self.handle is not None:
            (lib().cref_g1_bases_free if self.group == "g1" else lib().cref_g2_bases_free)(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass
