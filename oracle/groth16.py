"""bellman::groth16 generator / prover / verifier restated, generic over an engine.

ORACLE -- test infrastructure only (see oracle/__init__.py).

Follows:
  * groth16/generator.rs:67-170 (KeypairAssembly), :173-482 (generate_parameters)
  * groth16/prover.rs:45-82 (eval), :84-190 (ProvingAssignment), :205-364 (create_proof)
  * groth16/mod.rs:395-482 (ParameterSource for &Parameters), :27-53 (Proof)
  * groth16/verifier.rs:35-66 (verify_proof) -- via Engine.pairing_check
An engine is any object with: Fr (PrimeField), G1, G2 (group objects with the
oracle.curve.Curve interface) and `pairing_product_is_one(pairs)`.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Callable, List, Optional, Tuple

from .domain import EvaluationDomain
from .multiexp import UnexpectedIdentity, multiexp, SynthesisError


class UnconstrainedVariable(SynthesisError):
    pass


# ---------------------------------------------------------------------- minimal R1CS front end (bellman/src/lib.rs)
# Variable = ("in", i) | ("aux", i);  LinearCombination = list of (Variable, coeff:int)
ONE = ("in", 0)


class Circuit:
    def synthesize(self, cs):  # pragma: no cover - interface
        raise NotImplementedError


class KeypairAssembly:
    """generator.rs:67-170"""

    def __init__(self, F):
        self.F = F
        self.num_inputs = 0
        self.num_aux = 0
        self.num_constraints = 0
        self.at_inputs, self.bt_inputs, self.ct_inputs = [], [], []
        self.at_aux, self.bt_aux, self.ct_aux = [], [], []

    def alloc(self, f=None):
        i = self.num_aux
        self.num_aux += 1
        self.at_aux.append([]); self.bt_aux.append([]); self.ct_aux.append([])
        return ("aux", i)

    def alloc_input(self, f=None):
        i = self.num_inputs
        self.num_inputs += 1
        self.at_inputs.append([]); self.bt_inputs.append([]); self.ct_inputs.append([])
        return ("in", i)

    def enforce(self, a, b, c):
        def ev(lc, inputs, aux, n):
            for (kind, idx), coeff in lc:
                (inputs if kind == "in" else aux)[idx].append((coeff % self.F.p, n))
        n = self.num_constraints
        ev(a, self.at_inputs, self.at_aux, n)
        ev(b, self.bt_inputs, self.bt_aux, n)
        ev(c, self.ct_inputs, self.ct_aux, n)
        self.num_constraints += 1


@dataclass
class VerifyingKey:
    alpha_g1: object
    beta_g1: object
    beta_g2: object
    gamma_g2: object
    delta_g1: object
    delta_g2: object
    ic: list


@dataclass
class Parameters:
    """groth16/mod.rs:215-238"""
    vk: VerifyingKey
    h: list
    l: list
    a: list
    b_g1: list
    b_g2: list


@dataclass
class Proof:
    a: object
    b: object
    c: object


def generate_parameters(E, circuit, g1, g2, alpha, beta, gamma, delta, tau):
    """generator.rs:173-482.  g1, g2 are affine generators."""
    F = E.Fr
    asm = KeypairAssembly(F)
    asm.alloc_input()
    circuit.synthesize(asm)
    for i in range(asm.num_inputs):
        asm.enforce([(("in", i), 1)], [], [])

    dom = EvaluationDomain(F, [0] * asm.num_constraints)
    m = len(dom.coeffs)
    gamma_inverse = F.inv(gamma)
    delta_inverse = F.inv(delta)
    if gamma_inverse is None or delta_inverse is None:
        raise UnexpectedIdentity()
    # powers of tau
    cur = 1
    for i in range(m):
        dom.coeffs[i] = cur
        cur = F.mul(cur, tau)
    coeff = F.mul(dom.z(tau), delta_inverse)
    h = [E.G1.into_affine(E.G1.mul(g1, F.mul(dom.coeffs[i], coeff))) for i in range(m - 1)]
    dom.ifft()
    lag = dom.into_coeffs()

    def eval_at_tau(p):
        acc = 0
        for coeff_, index in p:
            acc = F.add(acc, F.mul(lag[index], coeff_))
        return acc

    def ev(at_l, bt_l, ct_l, inv):
        a, b1, b2, ext = [], [], [], []
        for at_p, bt_p, ct_p in zip(at_l, bt_l, ct_l):
            at = eval_at_tau(at_p)
            bt = eval_at_tau(bt_p)
            ct = eval_at_tau(ct_p)
            a.append(E.G1.mul(g1, at) if at != 0 else E.G1.zero())
            b1.append(E.G1.mul(g1, bt) if bt != 0 else E.G1.zero())
            b2.append(E.G2.mul(g2, bt) if bt != 0 else E.G2.zero())
            e = F.mul(F.add(F.add(F.mul(at, beta), F.mul(bt, alpha)), ct), inv)
            ext.append(E.G1.mul(g1, e))
        return a, b1, b2, ext

    a_in, b1_in, b2_in, ic = ev(asm.at_inputs, asm.bt_inputs, asm.ct_inputs, gamma_inverse)
    a_aux, b1_aux, b2_aux, l = ev(asm.at_aux, asm.bt_aux, asm.ct_aux, delta_inverse)
    for e in l:
        if E.G1.is_zero(e):
            raise UnconstrainedVariable()
    vk = VerifyingKey(
        alpha_g1=E.G1.into_affine(E.G1.mul(g1, alpha)),
        beta_g1=E.G1.into_affine(E.G1.mul(g1, beta)),
        beta_g2=E.G2.into_affine(E.G2.mul(g2, beta)),
        gamma_g2=E.G2.into_affine(E.G2.mul(g2, gamma)),
        delta_g1=E.G1.into_affine(E.G1.mul(g1, delta)),
        delta_g2=E.G2.into_affine(E.G2.mul(g2, delta)),
        ic=[E.G1.into_affine(e) for e in ic],
    )
    filt = lambda G, v: [G.into_affine(e) for e in v if not G.is_zero(e)]
    return Parameters(
        vk=vk,
        h=h,
        l=[E.G1.into_affine(e) for e in l],
        a=filt(E.G1, a_in + a_aux),
        b_g1=filt(E.G1, b1_in + b1_aux),
        b_g2=filt(E.G2, b2_in + b2_aux),
    ), asm


class ProvingAssignment:
    """prover.rs:84-190"""

    def __init__(self, F):
        self.F = F
        self.a_aux_density: List[bool] = []
        self.b_input_density: List[bool] = []
        self.b_aux_density: List[bool] = []
        self.a: List[int] = []
        self.b: List[int] = []
        self.c: List[int] = []
        self.input_assignment: List[int] = []
        self.aux_assignment: List[int] = []

    def alloc(self, f):
        self.aux_assignment.append(f() % self.F.p)
        self.a_aux_density.append(False)
        self.b_aux_density.append(False)
        return ("aux", len(self.aux_assignment) - 1)

    def alloc_input(self, f):
        self.input_assignment.append(f() % self.F.p)
        self.b_input_density.append(False)
        return ("in", len(self.input_assignment) - 1)

    def _eval(self, lc, input_density, aux_density):
        """prover.rs:45-82"""
        F = self.F
        acc = 0
        for (kind, i), coeff in lc:
            if kind == "in":
                tmp = self.input_assignment[i]
                if input_density is not None:
                    input_density[i] = True
            else:
                tmp = self.aux_assignment[i]
                if aux_density is not None:
                    aux_density[i] = True
            acc = F.add(acc, F.mul(tmp, coeff % F.p))
        return acc

    def enforce(self, a, b, c):
        self.a.append(self._eval(a, None, self.a_aux_density))
        self.b.append(self._eval(b, self.b_input_density, self.b_aux_density))
        self.c.append(self._eval(c, None, None))


def synthesize_assignment(E, circuit):
    """prover.rs:212-234"""
    prover = ProvingAssignment(E.Fr)
    prover.alloc_input(lambda: 1)
    circuit.synthesize(prover)
    for i in range(len(prover.input_assignment)):
        prover.enforce([(("in", i), 1)], [], [])
    return prover


def prove_from_assignment(E, prover, params, r, s, msm=None, hpoly=None):
    """prover.rs:249-364.  `msm(G, bases, exponents, density, base_offset)` and
    `hpoly(F, a, b, c)` can be overridden (the tests plug the CUDA path in here and
    compare the resulting proof with the all-oracle one)."""
    from .domain import h_coefficients
    F = E.Fr
    msm = msm or (lambda G, bases, exps, density, off: multiexp(G, bases, exps, density, off, F.num_bits))
    hpoly = hpoly or h_coefficients
    vk = params.vk
    h_coeffs = hpoly(F, prover.a, prover.b, prover.c)
    h = msm(E.G1, params.h, h_coeffs, None, 0)
    inputs = prover.input_assignment
    aux = prover.aux_assignment
    l = msm(E.G1, params.l, aux, None, 0)
    a_aux_total = sum(prover.a_aux_density)
    # get_a(num_inputs, _), groth16/mod.rs:456-463
    a_inputs = msm(E.G1, params.a, inputs, None, 0)
    a_aux = msm(E.G1, params.a, aux, prover.a_aux_density, len(inputs))
    b_in_total = sum(prover.b_input_density)
    b_g1_inputs = msm(E.G1, params.b_g1, inputs, prover.b_input_density, 0)
    b_g1_aux = msm(E.G1, params.b_g1, aux, prover.b_aux_density, b_in_total)
    b_g2_inputs = msm(E.G2, params.b_g2, inputs, prover.b_input_density, 0)
    b_g2_aux = msm(E.G2, params.b_g2, aux, prover.b_aux_density, b_in_total)
    if E.G1.affine_is_zero(vk.delta_g1) or E.G2.affine_is_zero(vk.delta_g2):
        raise UnexpectedIdentity()
    G1, G2 = E.G1, E.G2
    g_a = G1.add_mixed(G1.mul(vk.delta_g1, r), vk.alpha_g1)
    g_b = G2.add_mixed(G2.mul(vk.delta_g2, s), vk.beta_g2)
    rs = F.mul(r, s)
    g_c = G1.mul(vk.delta_g1, rs)
    g_c = G1.add(g_c, G1.mul(vk.alpha_g1, s))
    g_c = G1.add(g_c, G1.mul(vk.beta_g1, r))
    a_answer = G1.add(a_inputs, a_aux)
    g_a = G1.add(g_a, a_answer)
    a_answer = G1.mul_proj(a_answer, s)
    g_c = G1.add(g_c, a_answer)
    b1_answer = G1.add(b_g1_inputs, b_g1_aux)
    b2_answer = G2.add(b_g2_inputs, b_g2_aux)
    g_b = G2.add(g_b, b2_answer)
    b1_answer = G1.mul_proj(b1_answer, r)
    g_c = G1.add(g_c, b1_answer)
    g_c = G1.add(g_c, h)
    g_c = G1.add(g_c, l)
    return Proof(a=G1.into_affine(g_a), b=G2.into_affine(g_b), c=G1.into_affine(g_c))


def create_proof(E, circuit, params, r, s, **kw):
    """prover.rs:205-364"""
    return prove_from_assignment(E, synthesize_assignment(E, circuit), params, r, s, **kw)


def verify_proof(E, vk, proof, public_inputs):
    """verifier.rs:35-66: e(A,B) = e(alpha,beta) * e(acc,gamma) * e(C,delta)."""
    if len(public_inputs) + 1 != len(vk.ic):
        raise SynthesisError("MalformedVerifyingKey")
    G1, G2 = E.G1, E.G2
    acc = G1.into_projective(vk.ic[0])
    for i, b in zip(public_inputs, vk.ic[1:]):
        acc = G1.add(acc, G1.mul(b, i))
    acc = G1.into_affine(acc)
    return E.pairing_product_is_one([
        (proof.a, proof.b),
        (G1.affine_negate(vk.alpha_g1), vk.beta_g2),
        (G1.affine_negate(acc), vk.gamma_g2),
        (G1.affine_negate(proof.c), vk.delta_g2),
    ])


def proof_bytes(proof):
    """Proof::write, groth16/mod.rs:43-53: 48 + 96 + 48 compressed bytes (BLS12-381 only)."""
    from .curve import G1, G2
    return G1.encode_compressed(proof.a) + G2.encode_compressed(proof.b) + G1.encode_compressed(proof.c)


def verifying_key_bytes(vk):
    """VerifyingKey::write, groth16/mod.rs:140-158 (BLS12-381 only): six uncompressed points, u32 BE count, ic"""
    from .curve import G1, G2
    out = G1.encode_uncompressed(vk.alpha_g1) + G1.encode_uncompressed(vk.beta_g1) + G2.encode_uncompressed(vk.beta_g2)
    out += G2.encode_uncompressed(vk.gamma_g2) + G1.encode_uncompressed(vk.delta_g1) + G2.encode_uncompressed(vk.delta_g2)
    out += len(vk.ic).to_bytes(4, "big")
    for p in vk.ic:
        out += G1.encode_uncompressed(p)
    return out


def parameters_bytes(params):
    """Parameters::write, groth16/mod.rs:252-285: vk, then h, l, a, b_g1 (G1) and b_g2 (G2), each a u32 BE count + uncompressed points"""
    from .curve import G1, G2
    out = verifying_key_bytes(params.vk)
    for G, v in ((G1, params.h), (G1, params.l), (G1, params.a), (G1, params.b_g1), (G2, params.b_g2)):
        out += len(v).to_bytes(4, "big")
        for p in v:
            out += G.encode_uncompressed(p)
    return out
