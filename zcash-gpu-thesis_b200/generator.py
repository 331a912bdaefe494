"""groth16::generate_parameters (bellman/src/groth16/generator.rs:173-482) with the numeric part on the GPU.

Circuit synthesis into a `KeypairAssembly` (generator.rs:57-170) stays on the host, as in the reference; everything after
it runs through the C ABI on device-resident vectors:

  powers of tau             generator.rs:244-261   b200zk_distribute_powers_dev on a vector of ones
  h query                   generator.rs:263-288   b200zk_fr_scale_dev by (tau^m - 1) / delta, b200zk_fixed_base_mul_dev (G1)
  Lagrange coefficients     generator.rs:292       b200zk_ntt_dev (ifft)
  A / B / C at tau          generator.rs:361-380   b200zk_fr_spmv_dev over the per-variable (coeff, constraint) terms
  beta*A + alpha*B + C      generator.rs:382-398   b200zk_field_vec_dev / b200zk_fr_scale_dev
  a, b_g1, b_g2, ic, l      generator.rs:384-409   b200zk_fixed_base_mul_dev (the reference's wNAF tables, generator.rs:218-240)
  filtering of the zero points of a / b_g1 / b_g2, UnconstrainedVariable      generator.rs:452-480 (host)

The result is the same `Parameters` the reference builds: the point vectors are canonical affine Montgomery limbs.
"""
from __future__ import annotations

import numpy as np

from . import _lib as L
from .bellman import (FR_MODULUS, Parameters, PolynomialDegreeTooLarge, SynthesisError, UnexpectedIdentity, _ptr, _raise, _u64,
                      fr_from_mont_limbs, fr_to_mont_limbs)


class UnconstrainedVariable(SynthesisError):
    """SynthesisError::UnconstrainedVariable (generator.rs:452-456): an auxiliary variable that appears in no constraint"""


class KeypairAssembly:
    """generator.rs:57-170: records, per variable, the (coefficient, constraint index) terms of the A, B and C polynomials.
    Variables are ("in", i) / ("aux", i); linear combinations are lists of (variable, coefficient)."""

    def __init__(self):
        self.num_inputs = self.num_aux = self.num_constraints = 0
        self.at_inputs, self.bt_inputs, self.ct_inputs = [], [], []
        self.at_aux, self.bt_aux, self.ct_aux = [], [], []

    def alloc(self, _value=None):
        self.num_aux += 1
        for v in (self.at_aux, self.bt_aux, self.ct_aux):
            v.append([])
        return ("aux", self.num_aux - 1)

    def alloc_input(self, _value=None):
        self.num_inputs += 1
        for v in (self.at_inputs, self.bt_inputs, self.ct_inputs):
            v.append([])
        return ("in", self.num_inputs - 1)

    def enforce(self, a, b, c):
        for lc, ins, aux in ((a, self.at_inputs, self.at_aux), (b, self.bt_inputs, self.bt_aux), (c, self.ct_inputs, self.ct_aux)):
            for (kind, idx), coeff in lc:  # generator.rs:134-151 eval
                (ins if kind == "in" else aux)[idx].append((coeff % FR_MODULUS, self.num_constraints))
        self.num_constraints += 1


class GeneratedParameters:
    """Host copy of groth16::Parameters + VerifyingKey (groth16/mod.rs:100-126, 215-238): (n, 12) / (n, 24) uint64 arrays of
    affine Montgomery limbs.  `to_device` uploads it as the prover's `Parameters`."""

    def __init__(self, **kw):
        self.__dict__.update(kw)

    def to_device(self, worker, precompute=True) -> Parameters:
        p = Parameters(worker, self.h, self.l, self.a, self.b_g1, self.b_g2, self.alpha_g1, self.beta_g1, self.beta_g2, self.delta_g1, self.delta_g2)
        if precompute:
            for q in (p.h, p.l, p.a, p.b_g1, p.b_g2):
                q.precompute(0)
        return p


def _csr(rows):
    ptr = np.zeros(len(rows) + 1, dtype=np.uint32)
    cols, vals = [], []
    for i, row in enumerate(rows):
        for coeff, index in row:
            cols.append(index)
            vals.append(fr_to_mont_limbs(coeff))
        ptr[i + 1] = len(cols)
    col = np.array(cols, dtype=np.uint32) if cols else np.zeros(1, dtype=np.uint32)
    val = np.array(vals, dtype=np.uint64).reshape(-1, 4) if vals else np.zeros((1, 4), dtype=np.uint64)
    return ptr, col, val


def generate_parameters(worker, assembly, g1, g2, alpha, beta, gamma, delta, tau) -> GeneratedParameters:
    """generator.rs:173-482 after synthesis.  `assembly`: a KeypairAssembly that already holds the circuit *and* the input
    constraints of generator.rs:205-212; g1, g2: affine generators as 12 / 24 Montgomery u64; alpha..tau: ints in Fr."""
    lib, ctx, r = worker.lib, worker.ctx, FR_MODULUS

    def check(st):
        if st:
            _raise(worker, st)

    g1, g2 = _u64(g1), _u64(g2)
    m, log_m = 1, 0
    while m < assembly.num_constraints:  # EvaluationDomain::from_coeffs (domain.rs:48-81)
        m, log_m = m * 2, log_m + 1
        if log_m >= 32:
            raise PolynomialDegreeTooLarge()
    if gamma % r == 0 or delta % r == 0:
        raise UnexpectedIdentity("gamma / delta has no inverse (generator.rs:198-199)")
    gamma_inv, delta_inv = pow(gamma, r - 2, r), pow(delta, r - 2, r)
    n_in, n_aux = assembly.num_inputs, assembly.num_aux
    nv = n_in + n_aux
    bufs = []

    def alloc(nbytes):
        b = worker.alloc(max(nbytes, 32))
        bufs.append(b)
        return b

    def fixed_base(group, base, d_scalars_mont, n):
        """[s_i] base for Montgomery scalars on the device -> (n, 12|24) limbs and the zero flags on the host"""
        words = 12 if group == L.G1 else 24
        if n == 0:
            return np.zeros((0, words), np.uint64), np.zeros(0, np.uint8)
        d_rep, d_out, d_inf = alloc(n * 32), alloc(n * words * 8), alloc(n)
        check(lib.b200zk_field_vec_dev(ctx, L.FR, L.OP_INTO_REPR, d_scalars_mont, None, d_rep.ptr, n))  # Fr::into_repr, generator.rs:276
        check(lib.b200zk_fixed_base_mul_dev(ctx, group, _ptr(base), d_rep.ptr, n, 255, d_out.ptr, d_inf.ptr))
        return d_out.download(np.uint64, n * words).reshape(n, words), d_inf.download(np.uint8, n)

    try:
        # powers of tau (generator.rs:244-261)
        one = fr_to_mont_limbs(1)
        d_pow = alloc(m * 32).upload(np.tile(one, (m, 1)))
        check(lib.b200zk_distribute_powers_dev(ctx, d_pow.ptr, m, _ptr(fr_to_mont_limbs(tau))))
        # h[i] = g1^(tau^i * (tau^m - 1) / delta), i < m - 1 (generator.rs:263-288)
        z = np.zeros(4, np.uint64)
        check(lib.b200zk_domain_z(ctx, _ptr(fr_to_mont_limbs(tau)), log_m, _ptr(z)))
        coeff = fr_from_mont_limbs(z) * delta_inv % r
        d_h = alloc(m * 32)
        check(lib.b200zk_d2d(ctx, d_h.ptr, d_pow.ptr, m * 32))
        check(lib.b200zk_fr_scale_dev(ctx, d_h.ptr, m - 1, _ptr(fr_to_mont_limbs(coeff))))
        h, h_inf = fixed_base(L.G1, g1, d_h.ptr, m - 1)
        # Lagrange coefficients L_i(tau) (generator.rs:292-293)
        check(lib.b200zk_ntt_dev(ctx, d_pow.ptr, log_m, L.IFFT))
        # A, B, C at tau for every variable, inputs first (generator.rs:361-380)
        ev = []
        for ins, aux in ((assembly.at_inputs, assembly.at_aux), (assembly.bt_inputs, assembly.bt_aux), (assembly.ct_inputs, assembly.ct_aux)):
            ptr, col, val = _csr(list(ins) + list(aux))
            d_ptr, d_col, d_val, d_y = alloc(ptr.nbytes).upload(ptr), alloc(col.nbytes).upload(col), alloc(val.nbytes).upload(val), alloc(nv * 32)
            check(lib.b200zk_fr_spmv_dev(ctx, d_ptr.ptr, d_col.ptr, d_val.ptr, d_pow.ptr, nv, d_y.ptr))
            ev.append(d_y)
        d_at, d_bt, d_ct = ev
        # ext = (beta * at + alpha * bt + ct) * (gamma^-1 for inputs | delta^-1 for aux)  (generator.rs:382-398)
        d_t1, d_t2 = alloc(nv * 32), alloc(nv * 32)
        for dst, src, k in ((d_t1, d_at, beta), (d_t2, d_bt, alpha)):
            check(lib.b200zk_d2d(ctx, dst.ptr, src.ptr, nv * 32))
            check(lib.b200zk_fr_scale_dev(ctx, dst.ptr, nv, _ptr(fr_to_mont_limbs(k))))
        check(lib.b200zk_field_vec_dev(ctx, L.FR, L.OP_ADD, d_t1.ptr, d_t2.ptr, d_t1.ptr, nv))
        check(lib.b200zk_field_vec_dev(ctx, L.FR, L.OP_ADD, d_t1.ptr, d_ct.ptr, d_t1.ptr, nv))
        check(lib.b200zk_fr_scale_dev(ctx, d_t1.ptr, n_in, _ptr(fr_to_mont_limbs(gamma_inv))))
        check(lib.b200zk_fr_scale_dev(ctx, d_t1.ptr + n_in * 32, n_aux, _ptr(fr_to_mont_limbs(delta_inv))))
        a, a_inf = fixed_base(L.G1, g1, d_at.ptr, nv)
        b1, b1_inf = fixed_base(L.G1, g1, d_bt.ptr, nv)
        b2, b2_inf = fixed_base(L.G2, g2, d_bt.ptr, nv)
        ext, ext_inf = fixed_base(L.G1, g1, d_t1.ptr, nv)
        # the verifying key (generator.rs:462-470)
        d_k = alloc(4 * 32).upload(np.stack([fr_to_mont_limbs(x) for x in (alpha, beta, gamma, delta)]))
        vk1, _ = fixed_base(L.G1, g1, d_k.ptr, 4)
        vk2, _ = fixed_base(L.G2, g2, d_k.ptr, 4)
    finally:
        worker.sync()
        for b in bufs:
            b.free()
    if ext_inf[n_in:].any():  # generator.rs:452-456
        raise UnconstrainedVariable()
    keep = lambda pts, inf: np.ascontiguousarray(pts[inf == 0])  # generator.rs:478-480
    return GeneratedParameters(h=h, l=np.ascontiguousarray(ext[n_in:]), a=keep(a, a_inf), b_g1=keep(b1, b1_inf), b_g2=keep(b2, b2_inf),
                               ic=np.ascontiguousarray(ext[:n_in]), alpha_g1=vk1[0], beta_g1=vk1[1], delta_g1=vk1[3], beta_g2=vk2[1],
                               gamma_g2=vk2[2], delta_g2=vk2[3], h_infinity=h_inf, ic_infinity=ext_inf[:n_in])
