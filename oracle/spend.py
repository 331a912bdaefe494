"""create_proof after synthesis (bellman/src/groth16/prover.rs:249-364) on the CPU for LARGE circuits: the H block and the
eight multiexps by the C++ port (oracle/csrc/cref.cpp, the reference's own task structure), the assembly by the Python
restatement (oracle/curve.py).  Used to check Spend-shaped GPU proofs byte for byte and as the CPU baseline of proofs/s.

ORACLE -- test infrastructure only (see oracle/__init__.py).
"""
from __future__ import annotations

import time

import numpy as np

from . import cref
from .curve import G1, G2
from .fields import Fr


def _limbs_to_int(l):
    return sum(int(v) << (64 * i) for i, v in enumerate(l))


def _aff(G, limbs):
    limbs = [int(v) for v in limbs]
    h = len(limbs) // 2
    return (G.F.from_mont_limbs(limbs[:h]), G.F.from_mont_limbs(limbs[h:]), False)


class HostCrs:
    """groth16::Parameters on the host (groth16/mod.rs:215-238) in the reference's memory layout: five query vectors resident in
    the oracle + the VerifyingKey elements the prover reads."""

    def __init__(self, h, l, a, b_g1, b_g2, alpha_g1, beta_g1, beta_g2, delta_g1, delta_g2):
        self.h, self.l, self.a, self.b_g1 = (cref.ResidentBases.load("g1", v) for v in (h, l, a, b_g1))
        self.b_g2 = cref.ResidentBases.load("g2", b_g2)
        self.alpha_g1, self.beta_g1, self.delta_g1 = _aff(G1, alpha_g1), _aff(G1, beta_g1), _aff(G1, delta_g1)
        self.beta_g2, self.delta_g2 = _aff(G2, beta_g2), _aff(G2, delta_g2)


def multiexp_phase(crs: HostCrs, asg: dict, threads: int = 0):
    """prover.rs:256-318: the H block and the eight multiexps (C++ port, the reference's task structure).  Returns the eight
    Jacobian answers as a dict of oracle points."""
    n_con = asg["a"].shape[0]
    m = 1 << (n_con - 1).bit_length() if n_con > 1 else 1
    pad = lambda v: np.concatenate([v, np.zeros((m - v.shape[0], 4), dtype=np.uint64)])
    h_coeffs = cref.h_poly(pad(asg["a"]), pad(asg["b"]), pad(asg["c"]), threads)                # prover.rs:256-287
    inputs, aux = asg["inputs"], asg["aux"]
    n_in = inputs.shape[0]

    def msm(bases, exps, density, offset):
        st, jac = bases.multiexp(exps, density, offset, threads)
        assert st == 0, f"oracle multiexp failed with status {st}"
        return (G1 if bases.group == "g1" else G2).jacobian_from_limbs([int(v) for v in jac])

    b_in_total = int(np.count_nonzero(asg["b_input_density"]))                                   # :302-305
    return dict(
        h=msm(crs.h, h_coeffs, None, 0),                                                          # :289
        l=msm(crs.l, aux, None, 0),                                                               # :292
        a_inputs=msm(crs.a, inputs, None, 0),                                                     # :296-300, get_a(num_inputs, _)
        a_aux=msm(crs.a, aux, asg["a_aux_density"], n_in),
        b_g1_inputs=msm(crs.b_g1, inputs, asg["b_input_density"], 0),                             # :307-312
        b_g1_aux=msm(crs.b_g1, aux, asg["b_aux_density"], b_in_total),
        b_g2_inputs=msm(crs.b_g2, inputs, asg["b_input_density"], 0),                             # :314-318
        b_g2_aux=msm(crs.b_g2, aux, asg["b_aux_density"], b_in_total),
    )


def assemble(crs: HostCrs, q: dict, r: int, s: int) -> bytes:
    """prover.rs:326-363 + Proof::write (groth16/mod.rs:43-53)"""
    F = Fr
    g_a = G1.add_mixed(G1.mul(crs.delta_g1, r), crs.alpha_g1)
    g_b = G2.add_mixed(G2.mul(crs.delta_g2, s), crs.beta_g2)
    g_c = G1.mul(crs.delta_g1, F.mul(r, s))
    g_c = G1.add(g_c, G1.mul(crs.alpha_g1, s))
    g_c = G1.add(g_c, G1.mul(crs.beta_g1, r))
    a_answer = G1.add(q["a_inputs"], q["a_aux"])
    g_a = G1.add(g_a, a_answer)
    g_c = G1.add(g_c, G1.mul_proj(a_answer, s))
    b1_answer = G1.add(q["b_g1_inputs"], q["b_g1_aux"])
    g_b = G2.add(g_b, G2.add(q["b_g2_inputs"], q["b_g2_aux"]))
    g_c = G1.add(g_c, G1.mul_proj(b1_answer, r))
    g_c = G1.add(g_c, q["h"])
    g_c = G1.add(g_c, q["l"])
    return G1.encode_compressed(G1.into_affine(g_a)) + G2.encode_compressed(G2.into_affine(g_b)) + G1.encode_compressed(G1.into_affine(g_c))


def prove(crs: HostCrs, asg: dict, r: int, s: int, threads: int = 0):
    """prover.rs:249-364 for an already synthesized assignment (arrays as tools/synthetic.spend_assignment makes them).
    Returns (192 proof bytes, seconds spent in the H block + multiexps, seconds spent in the assembly)."""
    t0 = time.perf_counter()
    q = multiexp_phase(crs, asg, threads)
    t1 = time.perf_counter()
    out = assemble(crs, q, r, s)
    return out, t1 - t0, time.perf_counter() - t1
