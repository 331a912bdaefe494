"""Host-side mirror of the reference's operator interface for the prover hot path, over the C ABI.

Same names, argument meaning and error behaviour as the reference so the parity tests read like the
reference's own tests:

  bellman::multicore::Worker                      (bellman/src/multicore.rs:13-49)   -> Worker
  bellman::multiexp::{multiexp, FullDensity,
      DensityTracker, SourceBuilder}              (bellman/src/multiexp.rs:19-335)   -> multiexp, FullDensity, DensityTracker, Bases
  bellman::domain::EvaluationDomain               (bellman/src/domain.rs:26-189)     -> EvaluationDomain
  bellman::SynthesisError                         (bellman/src/lib.rs:171-188)       -> SynthesisError & subclasses

All arrays are numpy uint64 in the reference's memory layout (little-endian limbs; Montgomery field elements,
canonical FrRepr exponents).  All arithmetic happens on the GPU inside libb200zk.so; this file only moves
buffers and maps status codes to exceptions.  It never imports `oracle/`.
"""
from __future__ import annotations

import ctypes as C
import math

import numpy as np

from . import _lib as L

FR_MODULUS = 0x73eda753299d7d483339d80809a1d80553bda402fffe5bfeffffffff00000001
FR_S = 32  # fr.rs:47
_FR_R = (1 << 256) % FR_MODULUS
_FR_RINV = pow(_FR_R, -1, FR_MODULUS)


class SynthesisError(Exception):
    """bellman/src/lib.rs:171-188"""


class UnexpectedIdentity(SynthesisError):
    pass


class IoError(SynthesisError):
    """SynthesisError::IoError(UnexpectedEof): "expected more bases from source" (multiexp.rs:44-46)"""


class PolynomialDegreeTooLarge(SynthesisError):
    pass


class CudaError(RuntimeError):
    pass


class GroupDecodingError(ValueError):
    """pairing::GroupDecodingError (pairing/src/lib.rs:512-530), raised by the wire-format entry points"""


def _raise(worker, st):
    msg = worker.last_error() if worker is not None else ""
    if st == L.ERR_UNEXPECTED_IDENTITY:
        raise UnexpectedIdentity(msg)
    if st == L.ERR_UNEXPECTED_EOF:
        raise IoError(msg or "expected more bases from source")
    if st == L.ERR_DEGREE_TOO_LARGE:
        raise PolynomialDegreeTooLarge(msg)
    if st == L.ERR_DECODE:
        raise GroupDecodingError(msg)
    if st == L.ERR_BAD_ARG:
        raise ValueError(msg)
    raise CudaError(f"b200zk status {st}: {msg}")


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _u64(a, shape_last=None):
    a = np.ascontiguousarray(a, dtype=np.uint64)
    if shape_last is not None:
        a = a.reshape(-1, shape_last)
    return a


def fr_to_mont_limbs(x: int) -> np.ndarray:
    m = (x * _FR_R) % FR_MODULUS
    return np.array([(m >> (64 * i)) & 0xFFFFFFFFFFFFFFFF for i in range(4)], dtype=np.uint64)


def fr_from_mont_limbs(l) -> int:
    m = sum(int(v) << (64 * i) for i, v in enumerate(l))
    return (m * _FR_RINV) % FR_MODULUS


class DeviceBuffer:
    """A cudaMalloc'd region owned by a Worker (operands stay resident in HBM between calls)."""

    def __init__(self, worker, nbytes):
        self.worker = worker
        self.nbytes = int(nbytes)
        p = C.c_void_p()
        st = worker.lib.b200zk_dev_alloc(worker.ctx, self.nbytes, C.byref(p))
        if st:
            _raise(worker, st)
        self.ptr = p.value

    def upload(self, arr):
        arr = np.ascontiguousarray(arr)
        assert arr.nbytes <= self.nbytes
        st = self.worker.lib.b200zk_h2d(self.worker.ctx, self.ptr, _ptr(arr), arr.nbytes)
        if st:
            _raise(self.worker, st)
        self.worker.sync()  # the source array may be pageable / temporary
        return self

    def download(self, dtype, count):
        out = np.empty(count, dtype=dtype)
        st = self.worker.lib.b200zk_d2h(self.worker.ctx, _ptr(out), self.ptr, out.nbytes)
        if st:
            _raise(self.worker, st)
        return out

    def free(self):
        if self.ptr is not None and self.worker.ctx is not None:
            self.worker.lib.b200zk_dev_free(self.worker.ctx, self.ptr)
        self.ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class Worker:
    """bellman::multicore::Worker.  Here: one GPU + one CUDA stream (a b200zk context)."""

    def __init__(self, device: int = 0):
        self.lib = L.load()
        ctx = C.c_void_p()
        st = self.lib.b200zk_init(device, C.byref(ctx))
        if st:
            raise CudaError(f"b200zk_init(device={device}) failed with status {st}: no usable CUDA device (there is no CPU fallback)")
        self.ctx = ctx
        self.device = device
        self._pinned = []

    def last_error(self):
        return (self.lib.b200zk_last_error(self.ctx) or b"").decode()

    def sync(self):
        st = self.lib.b200zk_sync(self.ctx)
        if st:
            _raise(self, st)

    def sm_count(self):
        return self.lib.b200zk_sm_count(self.ctx)

    def alloc(self, nbytes):
        return DeviceBuffer(self, nbytes)

    def to_device(self, arr):
        arr = np.ascontiguousarray(arr)
        return DeviceBuffer(self, max(arr.nbytes, 1)).upload(arr)

    def timer_start(self):
        st = self.lib.b200zk_timer_start(self.ctx)
        if st:
            _raise(self, st)

    def timer_stop(self) -> float:
        ms = C.c_float()
        st = self.lib.b200zk_timer_stop(self.ctx, C.byref(ms))
        if st:
            _raise(self, st)
        return ms.value

    def set_msm_window(self, c):
        st = self.lib.b200zk_set_msm_window(self.ctx, c)
        if st:
            _raise(self, st)

    def microbench(self, kind, iters=2000):
        out = C.c_double()
        st = self.lib.b200zk_microbench(self.ctx, kind, iters, C.byref(out))
        if st:
            _raise(self, st)
        return out.value

    def pinned_array(self, shape, dtype=np.uint64):
        """A numpy array over page-locked host memory (b200zk_host_alloc_pinned): copies from it are true DMA transfers that
        overlap kernels, where a pageable Vec is staged through the driver.  Freed with the worker (or `free_pinned`)."""
        dt = np.dtype(dtype)
        count = int(np.prod(shape))
        hp = C.c_void_p()
        st = self.lib.b200zk_host_alloc_pinned(max(1, count * dt.itemsize), C.byref(hp))
        if st:
            _raise(self, st)
        self._pinned.append(hp)
        buf = (C.c_uint8 * (count * dt.itemsize)).from_address(hp.value)
        return np.frombuffer(buf, dtype=dt, count=count).reshape(shape)

    def pinned_copy(self, arr):
        arr = np.ascontiguousarray(arr)
        out = self.pinned_array(arr.shape, arr.dtype)
        out[...] = arr
        return out

    def close(self):
        if self.ctx is not None:
            for hp in self._pinned:
                self.lib.b200zk_host_free_pinned(hp)
            self._pinned = []
            self.lib.b200zk_destroy(self.ctx)
            self.ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


# ------------------------------------------------------------------------------------------------------ multiexp
class FullDensity:
    """multiexp.rs:78-97"""

    def get_query_size(self):
        return None


class DensityTracker:
    """multiexp.rs:99-138 (bit-vec storage expanded to one byte per exponent at the boundary)"""

    def __init__(self, bits=None):
        self.bv = [] if bits is None else [bool(b) for b in bits]
        self.total_density = sum(self.bv)

    def add_element(self):
        self.bv.append(False)

    def inc(self, idx):
        if not self.bv[idx]:
            self.bv[idx] = True
            self.total_density += 1

    def get_total_density(self):
        return self.total_density

    def get_query_size(self):
        return len(self.bv)

    def as_bytes(self):
        return np.array(self.bv, dtype=np.uint8)


class Bases:
    """The base vector of a SourceBuilder `(Arc<Vec<G>>, usize)` kept resident in HBM (multiexp.rs:34-68)."""

    def __init__(self, worker: Worker, group: int, xy, infinity=None, stride=None):
        self.worker = worker
        self.group = group
        width = 12 if group == L.G1 else 24
        xy = _u64(xy, width)
        self.n = xy.shape[0]
        inf = None if infinity is None else np.ascontiguousarray(infinity, dtype=np.uint8)
        h = C.c_void_p()
        st = worker.lib.b200zk_bases_upload(worker.ctx, group, _ptr(xy), self.n, width * 8, _ptr(inf), 1, C.byref(h))
        if st:
            _raise(worker, st)
        self.handle = h

    @classmethod
    def from_device(cls, worker, group, dbuf, n, dinf=None):
        self = cls.__new__(cls)
        self.worker, self.group, self.n = worker, group, n
        h = C.c_void_p()
        st = worker.lib.b200zk_bases_from_device(worker.ctx, group, dbuf.ptr, n, None if dinf is None else dinf.ptr, C.byref(h))
        if st:
            _raise(worker, st)
        self.handle = h
        return self

    @classmethod
    def read(cls, worker, group, data: bytes, checked=False, allow_infinity=False):
        """The point vectors of Parameters::read (groth16/mod.rs:287-382): uncompressed big-endian points decoded on the device."""
        pb = 96 if group == L.G1 else 192
        assert len(data) % pb == 0
        self = cls.__new__(cls)
        self.worker, self.group, self.n = worker, group, len(data) // pb
        buf = np.frombuffer(data, dtype=np.uint8)
        h = C.c_void_p()
        st = worker.lib.b200zk_bases_upload_encoded(worker.ctx, group, _ptr(buf), self.n, int(checked), int(allow_infinity), C.byref(h))
        if st:
            _raise(worker, st)
        self.handle = h
        return self

    def __len__(self):
        return self.n

    def precompute(self, window_bits=0):
        """Build the 2^(c w) * P table in HBM (one-time, at CRS load); multiexps on these bases then share one bucket set."""
        st = self.worker.lib.b200zk_bases_precompute(self.worker.ctx, self.handle, window_bits)
        if st:
            _raise(self.worker, st)
        return self

    def free(self):
        if getattr(self, "handle", None) is not None and self.worker.ctx is not None:
            self.worker.lib.b200zk_bases_free(self.handle)
        self.handle = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def window_size_reference(n: int) -> int:
    """The reference's window formula (multiexp.rs:296-300); used for the algorithmic-work accounting."""
    return 3 if n < 32 else int(math.ceil(math.log(float(n))))


def multiexp(pool: Worker, bases, density_map, exponents):
    """bellman::multiexp::multiexp (multiexp.rs:285-335).

    bases: a `Bases` or a `(Bases, offset)` SourceBuilder tuple; density_map: FullDensity or DensityTracker;
    exponents: (n, 4) uint64 canonical FrRepr.  Returns the Jacobian result as 18 (G1) / 36 (G2) uint64.
    Raises UnexpectedIdentity / IoError exactly where the reference's Source would (multiexp.rs:42-68).
    """
    if isinstance(bases, tuple):
        bases, offset = bases
    else:
        offset = 0
    exponents = _u64(exponents, 4)
    n = exponents.shape[0]
    density = None
    if density_map is not None and density_map.get_query_size() is not None:
        # multiexp.rs:302-307: the query size must match the number of exponents
        assert density_map.get_query_size() == n
        density = density_map.as_bytes()
    out = np.zeros(18 if bases.group == L.G1 else 36, dtype=np.uint64)
    if isinstance(pool, MultiWorker):  # one process, several GPUs: `bases` is a ShardedBases
        st = pool.lib.b200zk_multi_multiexp(pool.handle, bases.handle, offset, _ptr(exponents), n, _ptr(density), _ptr(out))
    else:
        st = pool.lib.b200zk_multiexp(pool.ctx, bases.handle, offset, _ptr(exponents), n, _ptr(density), _ptr(out))
    if st:
        _raise(pool, st)
    return out


def multiexp_batch(pool: Worker, bases, density_maps, exponents):
    """K multiexps over the same bases in one pipeline (b200zk_multiexp_batch_dev): `exponents` is (K, n, 4) canonical FrRepr,
    `density_maps` None (FullDensity for all) or a sequence of K DensityTracker / byte arrays.  Returns the K Jacobian
    results, shape (K, 18) or (K, 36); raises for the first multiexp whose Source would have failed."""
    if isinstance(bases, tuple):
        bases, offset = bases
    else:
        offset = 0
    exponents = np.ascontiguousarray(exponents, dtype=np.uint64)
    assert exponents.ndim == 3 and exponents.shape[2] == 4
    K, n = exponents.shape[0], exponents.shape[1]
    words = 18 if bases.group == L.G1 else 36
    if K == 0:
        return np.zeros((0, words), dtype=np.uint64)
    d_exp = pool.to_device(exponents)
    d_den = None
    if density_maps is not None:
        rows = [d.as_bytes() if isinstance(d, DensityTracker) else np.ascontiguousarray(d, dtype=np.uint8) for d in density_maps]
        assert len(rows) == K and all(r.shape[0] == n for r in rows)  # multiexp.rs:302-307
        d_den = pool.to_device(np.stack(rows) if n else np.zeros((K, 1), np.uint8))
    d_out, d_st = pool.alloc(K * words * 8), pool.alloc(K * 4)
    st = pool.lib.b200zk_multiexp_batch_dev(pool.ctx, bases.handle, offset, d_exp.ptr, n, n, d_den.ptr if d_den else None, n, K, d_out.ptr, d_st.ptr)
    if st:
        _raise(pool, st)
    out = d_out.download(np.uint64, K * words).reshape(K, words)
    codes = d_st.download(np.uint32, K)
    for x in (d_exp, d_den, d_out, d_st):
        if x is not None:
            x.free()
    for k in range(K):
        if codes[k] == L.ERR_UNEXPECTED_IDENTITY:
            raise UnexpectedIdentity(f"multiexp {k} of the batch consumed a base at infinity")
        if codes[k] == L.ERR_UNEXPECTED_EOF:
            raise IoError(f"multiexp {k} of the batch ran out of bases (UnexpectedEof)")
    return out


class MultiexpFuture:
    """The `Box<Future<Item = G::Projective, Error = SynthesisError>>` that multiexp returns (multiexp.rs:285-295)."""

    def __init__(self, pool, job, group, keep):
        self.pool, self.job, self.group, self.keep = pool, job, group, keep

    def wait(self):
        out = np.zeros(18 if self.group == L.G1 else 36, dtype=np.uint64)
        st = self.pool.lib.b200zk_job_wait(self.job, _ptr(out))
        self.job, self.keep = None, None
        if st:
            _raise(self.pool, st)
        return out


def multiexp_async(pool: Worker, bases, density_map, exponents) -> MultiexpFuture:
    """multiexp() as the reference uses it: returns immediately with a future; several may be in flight per Worker
    (prover.rs:289-318).  `exponents` should live in pinned memory for the copy to overlap the previous job."""
    if isinstance(bases, tuple):
        bases, offset = bases
    else:
        offset = 0
    exponents = _u64(exponents, 4)
    n = exponents.shape[0]
    density = None
    if density_map is not None and density_map.get_query_size() is not None:
        assert density_map.get_query_size() == n
        density = density_map.as_bytes()
    job = C.c_void_p()
    if isinstance(pool, MultiWorker):
        st = pool.lib.b200zk_multi_multiexp_async(pool.handle, bases.handle, offset, _ptr(exponents), n, _ptr(density), C.byref(job))
        if st:
            _raise(pool, st)
        return _GroupFuture(pool, job, bases.group, (exponents, density))
    st = pool.lib.b200zk_multiexp_async(pool.ctx, bases.handle, offset, _ptr(exponents), n, _ptr(density), C.byref(job))
    if st:
        _raise(pool, st)
    return MultiexpFuture(pool, job, bases.group, (exponents, density))


def into_affine(pool: Worker, group: int, jacobian):
    """CurveProjective::into_affine (ec.rs:586-619) on the device. Returns (xy array (n, 12|24), infinity flags)."""
    w = 18 if group == L.G1 else 36
    jac = _u64(jacobian, w)
    n = jac.shape[0]
    out = np.zeros((n, 12 if group == L.G1 else 24), dtype=np.uint64)
    inf = np.zeros(n, dtype=np.uint8)
    st = pool.lib.b200zk_into_affine(pool.ctx, group, _ptr(jac), n, _ptr(out), _ptr(inf))
    if st:
        _raise(pool, st)
    return out, inf


# ------------------------------------------------------------------------------------------------------ one process, several GPUs
class MultiWorker:
    """A Worker over several GPUs driven by ONE host process (b200zk_init_multi): what the reference's product caller needs, since
    a single zcashd process issues all the multiexps of a proof (prover.rs:289-318).  `devices` may repeat an id (two shards on
    one GPU).  `worker(i)` is a plain Worker view of shard i's context (for NTTs or proofs on a chosen GPU)."""

    def __init__(self, devices):
        self.lib = L.load()
        self.devices = [int(d) for d in devices]
        arr = (C.c_int * len(self.devices))(*self.devices)
        h = C.c_void_p()
        st = self.lib.b200zk_init_multi(arr, len(self.devices), C.byref(h))
        if st:
            raise CudaError(f"b200zk_init_multi({self.devices}) failed with status {st}: no usable CUDA device (there is no CPU fallback)")
        self.handle = h

    def __len__(self):
        return len(self.devices)

    def last_error(self):
        return (self.lib.b200zk_group_last_error(self.handle) or b"").decode()

    def peer_access(self, i):
        return bool(self.lib.b200zk_group_peer_access(self.handle, i))

    def worker(self, i) -> Worker:
        w = Worker.__new__(Worker)
        w.lib, w.device, w._pinned = self.lib, self.devices[i], []
        w.ctx = C.c_void_p(self.lib.b200zk_group_ctx(self.handle, i))
        w.close = lambda: None  # owned by the group
        return w

    def close(self):
        if self.handle is not None:
            self.lib.b200zk_group_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class ShardedBases:
    """A base vector sharded by contiguous range over the GPUs of a MultiWorker (b200zk_multi_bases_upload)."""

    def __init__(self, pool: MultiWorker, group: int, xy, infinity=None):
        self.worker, self.group = pool, group
        width = 12 if group == L.G1 else 24
        xy = _u64(xy, width)
        self.n = xy.shape[0]
        inf = None if infinity is None else np.ascontiguousarray(infinity, dtype=np.uint8)
        h = C.c_void_p()
        st = pool.lib.b200zk_multi_bases_upload(pool.handle, group, _ptr(xy), self.n, width * 8, _ptr(inf), 1, C.byref(h))
        if st:
            _raise(pool, st)
        self.handle = h

    def __len__(self):
        return self.n

    def precompute(self, window_bits=0):
        st = self.worker.lib.b200zk_multi_bases_precompute(self.worker.handle, self.handle, window_bits)
        if st:
            _raise(self.worker, st)
        return self

    def free(self):
        if getattr(self, "handle", None) is not None and self.worker.handle is not None:
            self.worker.lib.b200zk_multi_bases_free(self.handle)
        self.handle = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class _GroupFuture:
    def __init__(self, pool, job, group, keep):
        self.pool, self.job, self.group, self.keep = pool, job, group, keep

    def wait(self):
        out = np.zeros(18 if self.group == L.G1 else 36, dtype=np.uint64)
        st = self.pool.lib.b200zk_multi_job_wait(self.job, _ptr(out))
        self.job, self.keep = None, None
        if st:
            _raise(self.pool, st)
        return out


def multi_plan(bounds, base_offset, density, n_exp):
    """b200zk_multi_plan (host logic only, needs no GPU): exponent split points and per-shard base cursors of a sharded multiexp"""
    lib = L.load()
    nd = len(bounds) - 1
    b = (C.c_size_t * (nd + 1))(*bounds)
    e_lo = (C.c_size_t * (nd + 1))()
    off = (C.c_size_t * nd)()
    d = None if density is None else np.ascontiguousarray(density, dtype=np.uint8)
    st = lib.b200zk_multi_plan(b, nd, base_offset, _ptr(d), n_exp, e_lo, off)
    if st:
        raise ValueError(f"b200zk_multi_plan status {st}")
    return list(e_lo), list(off)


# ------------------------------------------------------------------------------------------------------ domain
class EvaluationDomain:
    """bellman::domain::EvaluationDomain<E, Scalar<E>> with the coefficients resident on the GPU."""

    def __init__(self, worker, dbuf, m, exp):
        self.worker, self.buf, self.m, self.exp = worker, dbuf, m, exp

    @classmethod
    def from_coeffs(cls, worker: Worker, coeffs):
        """domain.rs:48-81: pad to m = 2^exp with zeros; PolynomialDegreeTooLarge when exp >= Fr::S."""
        coeffs = _u64(coeffs, 4)
        m, exp = 1, 0
        while m < coeffs.shape[0]:
            m *= 2
            exp += 1
            if exp >= FR_S:
                raise PolynomialDegreeTooLarge()
        padded = np.zeros((m, 4), dtype=np.uint64)
        padded[: coeffs.shape[0]] = coeffs
        return cls(worker, worker.to_device(padded), m, exp)

    def __len__(self):
        return self.m

    def into_coeffs(self):
        return self.buf.download(np.uint64, self.m * 4).reshape(self.m, 4)

    def as_ref(self):
        return self.into_coeffs()

    def _ntt(self, kind):
        st = self.worker.lib.b200zk_ntt_dev(self.worker.ctx, self.buf.ptr, self.exp, kind)
        if st:
            _raise(self.worker, st)

    def fft(self, worker=None):
        self._ntt(L.FFT)

    def ifft(self, worker=None):
        self._ntt(L.IFFT)

    def coset_fft(self, worker=None):
        self._ntt(L.COSET_FFT)

    def icoset_fft(self, worker=None):
        self._ntt(L.ICOSET_FFT)

    def distribute_powers(self, worker, g: int):
        gl = fr_to_mont_limbs(g % FR_MODULUS)
        st = self.worker.lib.b200zk_distribute_powers_dev(self.worker.ctx, self.buf.ptr, self.m, _ptr(gl))
        if st:
            _raise(self.worker, st)

    def z(self, tau: int) -> int:
        """domain.rs:136-141: tau^m - 1 (computed on the device)"""
        out = np.zeros(4, dtype=np.uint64)
        st = self.worker.lib.b200zk_domain_z(self.worker.ctx, _ptr(fr_to_mont_limbs(tau % FR_MODULUS)), self.exp, _ptr(out))
        if st:
            _raise(self.worker, st)
        return fr_from_mont_limbs(out)

    def divide_by_z_on_coset(self, worker=None):
        """domain.rs:146-159: every coefficient times 1 / z(g), g = multiplicative_generator()"""
        st = self.worker.lib.b200zk_divide_by_z_on_coset_dev(self.worker.ctx, self.buf.ptr, self.exp)
        if st:
            _raise(self.worker, st)

    def _vec(self, op, other):
        assert self.m == other.m  # domain.rs:163, 179
        st = self.worker.lib.b200zk_field_vec_dev(self.worker.ctx, L.FR, op, self.buf.ptr, other.buf.ptr, self.buf.ptr, self.m)
        if st:
            _raise(self.worker, st)

    def mul_assign(self, worker, other):
        self._vec(L.OP_MUL, other)

    def sub_assign(self, worker, other):
        self._vec(L.OP_SUB, other)


def h_poly(worker: Worker, a, b, c):
    """The H-polynomial block of create_proof (prover.rs:256-287), fused on the device.
    a, b, c: evaluation vectors (n, 4) Montgomery; returns (m - 1, 4) canonical FrRepr limbs."""
    a, b, c = _u64(a, 4), _u64(b, 4), _u64(c, 4)
    n = max(a.shape[0], b.shape[0], c.shape[0])
    m, exp = 1, 0
    while m < n:
        m *= 2
        exp += 1
        if exp >= FR_S:
            raise PolynomialDegreeTooLarge()

    def pad(v):
        p = np.zeros((m, 4), dtype=np.uint64)
        p[: v.shape[0]] = v
        return p

    out = np.zeros((max(m - 1, 0), 4), dtype=np.uint64)
    if isinstance(worker, MultiWorker):  # a, b, c on three GPUs, peer-to-peer combine on the first (b200zk_multi_h_poly)
        st = worker.lib.b200zk_multi_h_poly(worker.handle, _ptr(pad(a)), _ptr(pad(b)), _ptr(pad(c)), exp, _ptr(out))
    else:
        st = worker.lib.b200zk_h_poly(worker.ctx, _ptr(pad(a)), _ptr(pad(b)), _ptr(pad(c)), exp, _ptr(out))
    if st:
        _raise(worker, st)
    return out


def decode_points(worker, group, data: bytes, checked=False):
    """EncodedPoint::into_affine[_unchecked] for uncompressed encodings (ec.rs:686-752): (xy Montgomery limbs, infinity flags)"""
    pb = 96 if group == L.G1 else 192
    n = len(data) // pb
    buf = np.frombuffer(data, dtype=np.uint8)
    out = np.zeros((n, pb // 8), dtype=np.uint64)
    inf = np.zeros(n, dtype=np.uint8)
    st = worker.lib.b200zk_decode_points(worker.ctx, group, _ptr(buf), n, int(checked), _ptr(out), _ptr(inf))
    if st:
        _raise(worker, st)
    return out, inf


def encode_points(worker, group, xy, inf=None, compressed=False) -> bytes:
    """into_uncompressed / into_compressed (ec.rs:796-868, 2750-2830) on the device"""
    w = 12 if group == L.G1 else 24
    xy = _u64(xy, w)
    n = xy.shape[0]
    if inf is not None:
        inf = np.ascontiguousarray(inf, dtype=np.uint8)
    size = (w * 8) // (2 if compressed else 1)
    out = np.zeros(n * size, dtype=np.uint8)
    st = worker.lib.b200zk_encode_points(worker.ctx, group, _ptr(xy), _ptr(inf), n, int(compressed), _ptr(out))
    if st:
        _raise(worker, st)
    return out.tobytes()


def ntt_plan(log_m, large_from=20, sm_count=148, batch=1):
    """b200zk_ntt_plan (host logic only, needs no GPU): [(stages, columns_log), ...] per pass and whether the radix-4 kernels run"""
    lib = L.load()
    stages, cols = (C.c_uint32 * 8)(), (C.c_uint32 * 8)()
    radix4 = C.c_int(0)
    n = lib.b200zk_ntt_plan(log_m, large_from, sm_count, batch, stages, cols, C.byref(radix4))
    return [(int(stages[i]), int(cols[i])) for i in range(n)], bool(radix4.value)


# ------------------------------------------------------------------------------------------------------ test / bench helpers
def field_vec(worker, field, op, a, b=None):
    """element-wise Fr / Fq / Fq2 ops on host arrays; OP_MULSUB (Fq): rows of a are (p, q), rows of b are (r, s), out = p q - r s"""
    w = {L.FR: 4, L.FQ: 6, L.FQ2: 12, L.FQ2_PAIR: 12}[field]
    wi = 2 * w if op == L.OP_MULSUB else w  # (FQ2_PAIR MULSUB: rows of two Fq2 elements)
    a = _u64(a, wi)
    b = None if b is None else _u64(b, wi)
    out = np.empty((a.shape[0], w), dtype=np.uint64)
    st = worker.lib.b200zk_field_vec(worker.ctx, field, op, _ptr(a), _ptr(b), _ptr(out), a.shape[0])
    if st:
        _raise(worker, st)
    return out


def point_op(worker, group, op, a, b=None, b_inf=None):
    w = 18 if group == L.G1 else 36
    a = _u64(a, w)
    n = a.shape[0]
    if b is not None:
        b = _u64(b, w if op == L.POINT_ADD else (12 if group == L.G1 else 24))
    if b_inf is not None:
        b_inf = np.ascontiguousarray(b_inf, dtype=np.uint8)
    out = np.empty_like(a)
    st = worker.lib.b200zk_point_op(worker.ctx, group, op, _ptr(a), _ptr(b), _ptr(b_inf), _ptr(out), n)
    if st:
        _raise(worker, st)
    return out


def ntt_host(worker, coeffs, kind):
    """b200zk_ntt on a host array (copy in, transform, copy out)."""
    a = _u64(coeffs, 4).copy()
    m = a.shape[0]
    exp = m.bit_length() - 1
    assert 1 << exp == m
    st = worker.lib.b200zk_ntt(worker.ctx, _ptr(a), exp, kind)
    if st:
        _raise(worker, st)
    return a


def fixed_base_mul(worker, group, base_xy, scalars, scalar_bits=255):
    """out[i] = scalars[i] * base, computed and left on the device.  Returns (DeviceBuffer xy, DeviceBuffer inf, n)."""
    scalars = _u64(scalars, 4)
    n = scalars.shape[0]
    base_xy = _u64(base_xy)
    ds = worker.to_device(scalars)
    pb = 96 if group == L.G1 else 192
    dout = worker.alloc(max(n * pb, 1))
    dinf = worker.alloc(max(n, 1))
    st = worker.lib.b200zk_fixed_base_mul_dev(worker.ctx, group, _ptr(base_xy), ds.ptr, n, scalar_bits, dout.ptr, dinf.ptr)
    if st:
        _raise(worker, st)
    worker.sync()
    ds.free()
    return dout, dinf, n


# ------------------------------------------------------------------------------------------------------ groth16
FQ_MODULUS = 0x1a0111ea397fe69a4b1ba7b6434bacd764774b84f38512bf6730d2a0f6b0f6241eabfffeb153ffffb9feffffffffaaab


class Parameters:
    """groth16::Parameters (groth16/mod.rs:215-238) resident in HBM: the h, l, a, b_g1, b_g2 query vectors as `Bases`
    and the VerifyingKey elements the prover reads.  Acts as the ParameterSource of create_proof."""

    def __init__(self, worker, h, l, a, b_g1, b_g2, alpha_g1, beta_g1, beta_g2, delta_g1, delta_g2, vk_infinity=None):
        self.worker = worker
        mk = lambda g, v: v if isinstance(v, Bases) else Bases(worker, g, v)
        self.h, self.l, self.a, self.b_g1 = (mk(L.G1, v) for v in (h, l, a, b_g1))
        self.b_g2 = mk(L.G2, b_g2)
        vk = [_u64(v) for v in (alpha_g1, beta_g1, beta_g2, delta_g1, delta_g2)]
        inf = None if vk_infinity is None else np.ascontiguousarray(vk_infinity, dtype=np.uint8)
        hnd = C.c_void_p()
        st = worker.lib.b200zk_crs_create(worker.ctx, self.h.handle, self.l.handle, self.a.handle, self.b_g1.handle, self.b_g2.handle,
                                          _ptr(vk[0]), _ptr(vk[1]), _ptr(vk[2]), _ptr(vk[3]), _ptr(vk[4]), _ptr(inf), C.byref(hnd))
        if st:
            _raise(worker, st)
        self.handle = hnd

    @classmethod
    def read(cls, worker, data: bytes, checked: bool = False, precompute: bool = True):
        """Parameters::read (groth16/mod.rs:287-382): the proving key from its wire format, decoded on the device and kept
        resident.  `checked` = into_affine (curve + subgroup tests) for the query vectors; the vk is always checked.
        Raises GroupDecodingError for malformed or truncated input and for a point at infinity in ic / a query vector."""
        self = cls.__new__(cls)
        self.worker = worker
        buf = np.frombuffer(bytes(data), dtype=np.uint8)
        hnd = C.c_void_p()
        st = worker.lib.b200zk_parameters_read(worker.ctx, _ptr(buf), buf.shape[0], int(checked), C.byref(hnd))
        if st:
            _raise(worker, st)
        self.handle = hnd
        self.h = self.l = self.a = self.b_g1 = self.b_g2 = None  # owned by the CRS object
        if precompute:
            st = worker.lib.b200zk_crs_precompute(worker.ctx, hnd, 0)
            if st:
                _raise(worker, st)
        return self

    def write(self) -> bytes:
        """Parameters::write (groth16/mod.rs:252-285); only for parameters that came from `read` (they carry gamma_g2 and ic)."""
        size = self.worker.lib.b200zk_parameters_size(self.handle)
        if size == 0:
            raise ValueError("these parameters were not made by Parameters.read: gamma_g2 and ic are unknown")
        out = np.zeros(size, dtype=np.uint8)
        st = self.worker.lib.b200zk_parameters_write(self.worker.ctx, self.handle, _ptr(out), size)
        if st:
            _raise(self.worker, st)
        return out.tobytes()

    def query_sizes(self):
        sizes = (C.c_size_t * 5)()
        st = self.worker.lib.b200zk_crs_query_sizes(self.handle, sizes)
        if st:
            _raise(self.worker, st)
        return dict(zip(("h", "l", "a", "b_g1", "b_g2"), (int(v) for v in sizes)))

    def verifying_key(self):
        """groth16::VerifyingKey (mod.rs:100-126) as Montgomery limb arrays: dict with alpha_g1, beta_g1, beta_g2, gamma_g2,
        delta_g1, delta_g2 (each (xy, infinity)) and ic (n, 12)"""
        vk = np.zeros(108, dtype=np.uint64)
        inf = np.zeros(6, dtype=np.uint8)
        n = C.c_size_t()
        st = self.worker.lib.b200zk_crs_verifying_key(self.handle, _ptr(vk), _ptr(inf), None, C.byref(n))
        if st:
            raise ValueError("these parameters carry no full VerifyingKey (not made by Parameters.read)")
        ic = np.zeros((n.value, 12), dtype=np.uint64)
        self.worker.lib.b200zk_crs_verifying_key(self.handle, None, None, _ptr(ic), C.byref(n))
        off = dict(alpha_g1=(0, 12), beta_g1=(12, 24), beta_g2=(24, 48), gamma_g2=(48, 72), delta_g1=(72, 84), delta_g2=(84, 108))
        out = {k: (vk[a:b].copy(), bool(inf[i])) for i, (k, (a, b)) in enumerate(off.items())}
        out["ic"] = ic
        return out

    def free(self):
        if getattr(self, "handle", None) is not None and self.worker.ctx is not None:
            self.worker.lib.b200zk_crs_free(self.handle)
        self.handle = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class Proof:
    """groth16::Proof (groth16/mod.rs:27-53): affine a (G1), b (G2), c (G1) as Montgomery limbs + infinity flags."""

    def __init__(self, a, b, c, inf):
        self.a, self.b, self.c, self.inf = a, b, c, [bool(x) for x in inf]

    def write(self, worker) -> bytes:
        """Proof::write: 48 + 96 + 48 compressed big-endian bytes (groth16/mod.rs:43-53; ec.rs:839-868, 2801-2830), encoded on the device."""
        a = encode_points(worker, L.G1, self.a, [self.inf[0]], compressed=True)
        b = encode_points(worker, L.G2, self.b, [self.inf[1]], compressed=True)
        c = encode_points(worker, L.G1, self.c, [self.inf[2]], compressed=True)
        return a + b + c


def create_proof_from_assignment(worker, params: Parameters, a, b, c, input_assignment, aux_assignment, a_aux_density, b_input_density,
                                 b_aux_density, r, s) -> Proof:
    """groth16::create_proof (prover.rs:205-364) after circuit synthesis: a, b, c = ProvingAssignment's evaluation vectors
    ((n, 4) Montgomery), input/aux assignments as canonical FrRepr ((n, 4)), densities as DensityTracker or byte arrays,
    r, s as Python ints (E::Fr)."""
    a, b, c = _u64(a, 4), _u64(b, 4), _u64(c, 4)
    inputs, aux = _u64(input_assignment, 4), _u64(aux_assignment, 4)
    dens = []
    for d in (a_aux_density, b_input_density, b_aux_density):
        dens.append(d.as_bytes() if isinstance(d, DensityTracker) else np.ascontiguousarray(d, dtype=np.uint8))
    assert dens[0].shape[0] == aux.shape[0] and dens[1].shape[0] == inputs.shape[0] and dens[2].shape[0] == aux.shape[0]
    if b.shape[0] != a.shape[0] or c.shape[0] != a.shape[0] or a.shape[0] == 0 or inputs.shape[0] == 0:
        raise ValueError("a, b and c must hold one evaluation per constraint (equal, non-zero lengths) and there is at least the input ONE")
    rl = np.array([(r >> (64 * i)) & 0xFFFFFFFFFFFFFFFF for i in range(4)], dtype=np.uint64)
    sl = np.array([(s >> (64 * i)) & 0xFFFFFFFFFFFFFFFF for i in range(4)], dtype=np.uint64)
    pa, pb, pc = np.zeros(12, np.uint64), np.zeros(24, np.uint64), np.zeros(12, np.uint64)
    inf = np.zeros(3, np.uint8)
    st = worker.lib.b200zk_groth16_prove(worker.ctx, params.handle, _ptr(a), _ptr(b), _ptr(c), a.shape[0], _ptr(inputs), inputs.shape[0],
                                         _ptr(aux), aux.shape[0], _ptr(dens[0]), _ptr(dens[1]), _ptr(dens[2]), _ptr(rl), _ptr(sl),
                                         _ptr(pa), _ptr(pb), _ptr(pc), _ptr(inf))
    if st:
        _raise(worker, st)
    return Proof(pa, pb, pc, inf)


def create_proofs_from_assignments(worker, params: Parameters, assignments, lockstep=0):
    """A batch of create_proof calls over one CRS and one circuit (b200zk_groth16_prove_batch): `assignments` is a sequence of
    (a, b, c, input_assignment, aux_assignment, a_aux_density, b_input_density, b_aux_density, r, s) tuples with the meaning of
    create_proof_from_assignment's arguments; all of the same sizes.  Returns the list of Proofs, in order."""
    n = len(assignments)
    if n == 0:
        return []
    keep, rows = [], (L.ProveInput * n)()
    shape = None
    for i, (a, b, c, inputs, aux, da, dbi, dba, r, s) in enumerate(assignments):
        a, b, c, inputs, aux = _u64(a, 4), _u64(b, 4), _u64(c, 4), _u64(inputs, 4), _u64(aux, 4)
        dens = [d.as_bytes() if isinstance(d, DensityTracker) else np.ascontiguousarray(d, dtype=np.uint8) for d in (da, dbi, dba)]
        assert dens[0].shape[0] == aux.shape[0] and dens[1].shape[0] == inputs.shape[0] and dens[2].shape[0] == aux.shape[0]
        this = (a.shape[0], inputs.shape[0], aux.shape[0])
        if shape is None:
            shape = this
        elif this != shape or b.shape[0] != shape[0] or c.shape[0] != shape[0]:
            raise ValueError("the proofs of a batch must come from the same circuit (equal sizes)")
        rl = np.array([(r >> (64 * j)) & 0xFFFFFFFFFFFFFFFF for j in range(4)], dtype=np.uint64)
        sl = np.array([(s >> (64 * j)) & 0xFFFFFFFFFFFFFFFF for j in range(4)], dtype=np.uint64)
        arrs = (a, b, c, inputs, aux, dens[0], dens[1], dens[2], rl, sl)
        keep.append(arrs)
        for (name, _), arr in zip(L.ProveInput._fields_, arrs):
            setattr(rows[i], name, arr.ctypes.data)
    pa, pb, pc = np.zeros((n, 12), np.uint64), np.zeros((n, 24), np.uint64), np.zeros((n, 12), np.uint64)
    inf = np.zeros((n, 3), np.uint8)
    st = worker.lib.b200zk_groth16_prove_batch(worker.ctx, params.handle, C.cast(rows, C.c_void_p), n, shape[0], shape[1], shape[2], lockstep,
                                               _ptr(pa), _ptr(pb), _ptr(pc), _ptr(inf))
    if st:
        _raise(worker, st)
    return [Proof(pa[i], pb[i], pc[i], inf[i]) for i in range(n)]


def create_proof_bytes_from_assignments(worker, params: Parameters, assignments, lockstep=0):
    """The batch as the outer FFI sees it (rustzcash.rs:1375-1626 -> create_random_proof -> Proof::write): a sequence of assignment
    tuples in, the list of 192-byte proofs out (b200zk_groth16_prove_batch_bytes: proved in lock-step groups, encoded on the device)."""
    n = len(assignments)
    if n == 0:
        return []
    keep, rows = [], (L.ProveInput * n)()
    shape = None
    for i, (a, b, c, inputs, aux, da, dbi, dba, r, s) in enumerate(assignments):
        a, b, c, inputs, aux = _u64(a, 4), _u64(b, 4), _u64(c, 4), _u64(inputs, 4), _u64(aux, 4)
        dens = [d.as_bytes() if isinstance(d, DensityTracker) else np.ascontiguousarray(d, dtype=np.uint8) for d in (da, dbi, dba)]
        this = (a.shape[0], inputs.shape[0], aux.shape[0])
        if shape is None:
            shape = this
        if this != shape or b.shape[0] != shape[0] or c.shape[0] != shape[0] or dens[0].shape[0] != shape[2] or dens[1].shape[0] != shape[1] or dens[2].shape[0] != shape[2]:
            raise ValueError("the proofs of a batch must come from the same circuit (equal sizes)")
        rl = np.array([(r >> (64 * j)) & 0xFFFFFFFFFFFFFFFF for j in range(4)], dtype=np.uint64)
        sl = np.array([(s >> (64 * j)) & 0xFFFFFFFFFFFFFFFF for j in range(4)], dtype=np.uint64)
        arrs = (a, b, c, inputs, aux, dens[0], dens[1], dens[2], rl, sl)
        keep.append(arrs)
        for (name, _), arr in zip(L.ProveInput._fields_, arrs):
            setattr(rows[i], name, arr.ctypes.data)
    out = np.zeros(192 * n, dtype=np.uint8)
    st = worker.lib.b200zk_groth16_prove_batch_bytes(worker.ctx, params.handle, C.cast(rows, C.c_void_p), n, shape[0], shape[1], shape[2], lockstep, _ptr(out))
    if st:
        _raise(worker, st)
    return [out[192 * i:192 * (i + 1)].tobytes() for i in range(n)]


# ------------------------------------------------------------------------------------------------------ verifier
def pairing(worker, g1_xy, g2_xy, g1_inf=None, g2_inf=None):
    """Engine::pairing (pairing/src/lib.rs:86-96) for n independent pairs on the device: (n, 72) uint64, the Fq12 values as
    c0.c0.c0, c0.c0.c1, c0.c1.c0, ... Montgomery limbs (the order of the reference's known-answer literal)."""
    p, q = _u64(g1_xy, 12), _u64(g2_xy, 24)
    n = p.shape[0]
    assert q.shape[0] == n
    pi = None if g1_inf is None else np.ascontiguousarray(g1_inf, dtype=np.uint8)
    qi = None if g2_inf is None else np.ascontiguousarray(g2_inf, dtype=np.uint8)
    out = np.zeros((n, 72), dtype=np.uint64)
    st = worker.lib.b200zk_pairing(worker.ctx, _ptr(p), _ptr(pi), _ptr(q), _ptr(qi), n, _ptr(out))
    if st:
        _raise(worker, st)
    return out


class PreparedVerifyingKey:
    """groth16::PreparedVerifyingKey (groth16/mod.rs:384-393) made by prepare_verifying_key (verifier.rs:18-33), resident on the GPU"""

    def __init__(self, worker, alpha_g1, beta_g2, gamma_g2, delta_g2, ic):
        self.worker = worker
        ic = _u64(ic, 12)
        self.n_ic = ic.shape[0]
        h = C.c_void_p()
        st = worker.lib.b200zk_prepare_verifying_key(worker.ctx, _ptr(_u64(alpha_g1)), _ptr(_u64(beta_g2)), _ptr(_u64(gamma_g2)), _ptr(_u64(delta_g2)),
                                                     _ptr(ic), self.n_ic, C.byref(h))
        if st:
            _raise(worker, st)
        self.handle = h

    def free(self):
        if getattr(self, "handle", None) is not None and self.worker.ctx is not None:
            self.worker.lib.b200zk_pvk_free(self.handle)
        self.handle = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def prepare_verifying_key(worker, vk) -> PreparedVerifyingKey:
    """verifier.rs:18-33; `vk` is the dict Parameters.verifying_key() returns"""
    return PreparedVerifyingKey(worker, vk["alpha_g1"][0], vk["beta_g2"][0], vk["gamma_g2"][0], vk["delta_g2"][0], vk["ic"])


def verify_proofs(worker, pvk: PreparedVerifyingKey, proofs, public_inputs):
    """groth16::verify_proof (verifier.rs:35-66) for a batch on the device: `proofs` is a sequence of Proof, `public_inputs` an
    (n, n_inputs) sequence of ints (without the leading ONE).  Returns a list of bools.  Raises ValueError
    (MalformedVerifyingKey) when the number of inputs does not fit the key."""
    n = len(proofs)
    if n == 0:
        return []
    a = np.stack([_u64(p.a) for p in proofs]).reshape(n, 12)
    b = np.stack([_u64(p.b) for p in proofs]).reshape(n, 24)
    c = np.stack([_u64(p.c) for p in proofs]).reshape(n, 12)
    inf = np.array([[int(x) for x in p.inf] for p in proofs], dtype=np.uint8).reshape(n, 3)
    n_in = len(public_inputs[0]) if n else 0
    assert all(len(row) == n_in for row in public_inputs)
    ins = np.array([[(int(v) % FR_MODULUS >> (64 * j)) & 0xFFFFFFFFFFFFFFFF for v in row for j in range(4)] for row in public_inputs], dtype=np.uint64).reshape(n, max(n_in, 1) * 4 if n_in else 0)
    ok = np.zeros(n, dtype=np.uint8)
    st = worker.lib.b200zk_verify_proofs(worker.ctx, pvk.handle, _ptr(a), _ptr(b), _ptr(c), _ptr(inf), _ptr(ins) if n_in else None, n_in, n, _ptr(ok))
    if st:
        _raise(worker, st)
    return [bool(x) for x in ok]


def verify_proof(worker, pvk: PreparedVerifyingKey, proof, public_inputs) -> bool:
    return verify_proofs(worker, pvk, [proof], [list(public_inputs)])[0]


# ------------------------------------------------------------------------------------------- circuit-facing prover API
ONE = ("in", 0)  # ConstraintSystem::one() (bellman/src/lib.rs): the first input variable


class ProvingAssignment:
    """prover.rs:84-190: the ConstraintSystem a circuit is synthesized into by create_proof.  Collects the a / b / c evaluation
    vectors, the input / aux assignments and the three density trackers; host arithmetic (Python ints mod r), as in the
    reference synthesis is CPU work.  Variables are ("in", i) / ("aux", i); a linear combination is [(variable, coeff), ...]."""

    def __init__(self):
        self.a, self.b, self.c = [], [], []
        self.input_assignment, self.aux_assignment = [], []
        self.a_aux_density, self.b_input_density, self.b_aux_density = [], [], []

    def alloc(self, f):
        self.aux_assignment.append(f() % FR_MODULUS)
        self.a_aux_density.append(0)
        self.b_aux_density.append(0)
        return ("aux", len(self.aux_assignment) - 1)

    def alloc_input(self, f):
        self.input_assignment.append(f() % FR_MODULUS)
        self.b_input_density.append(0)
        return ("in", len(self.input_assignment) - 1)

    def _eval(self, lc, input_density, aux_density):  # prover.rs:45-82
        acc = 0
        for (kind, i), coeff in lc:
            if kind == "in":
                v = self.input_assignment[i]
                if input_density is not None:
                    input_density[i] = 1
            else:
                v = self.aux_assignment[i]
                if aux_density is not None:
                    aux_density[i] = 1
            acc += v * (coeff % FR_MODULUS)
        return acc % FR_MODULUS

    def enforce(self, a, b, c):  # prover.rs:153-186
        self.a.append(self._eval(a, None, self.a_aux_density))  # inputs have full density in the A query (prover.rs:161-166)
        self.b.append(self._eval(b, self.b_input_density, self.b_aux_density))
        self.c.append(self._eval(c, None, None))

    def as_tuple(self, r, s):
        mont = lambda v: np.array([fr_to_mont_limbs(x) for x in v], dtype=np.uint64).reshape(len(v), 4)
        rep = lambda v: np.array([[(x >> (64 * j)) & 0xFFFFFFFFFFFFFFFF for j in range(4)] for x in v], dtype=np.uint64).reshape(len(v), 4)
        u8 = lambda v: np.array(v, dtype=np.uint8)
        return (mont(self.a), mont(self.b), mont(self.c), rep(self.input_assignment), rep(self.aux_assignment), u8(self.a_aux_density),
                u8(self.b_input_density), u8(self.b_aux_density), r % FR_MODULUS, s % FR_MODULUS)


def synthesize(circuit) -> ProvingAssignment:
    """prover.rs:212-234: allocate ONE, synthesize the circuit, add the `input * 0 = 0` rows"""
    prover = ProvingAssignment()
    prover.alloc_input(lambda: 1)
    circuit.synthesize(prover)
    for i in range(len(prover.input_assignment)):
        prover.enforce([(("in", i), 1)], [], [])
    return prover


def create_proof(worker, circuit, params: Parameters, r: int, s: int) -> Proof:
    """groth16::create_proof (prover.rs:205-364): synthesis on the host, everything else in b200zk_groth16_prove"""
    return create_proof_from_assignment(worker, params, *synthesize(circuit).as_tuple(r, s))


def create_random_proof(worker, circuit, params: Parameters, rng) -> Proof:
    """groth16::create_random_proof (prover.rs:192-203): r, s <- rng (`rng.randrange(n)` like random.Random / SystemRandom)"""
    return create_proof(worker, circuit, params, rng.randrange(FR_MODULUS), rng.randrange(FR_MODULUS))


def create_proofs(worker, circuits, params: Parameters, rs, lockstep=0):
    """A run of create_proof calls over one CRS and one circuit shape, proved in lock-step groups (b200zk_groth16_prove_batch)"""
    return create_proofs_from_assignments(worker, params, [synthesize(c).as_tuple(r, s) for c, (r, s) in zip(circuits, rs)], lockstep)
