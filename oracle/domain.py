"""bellman::domain::EvaluationDomain restated (bellman/src/domain.rs:26-374).

ORACLE -- test infrastructure only (see oracle/__init__.py).

Generic over a prime field `F` (oracle.fields.PrimeField: BLS12-381 Fr, or the
64513-element field of the reference's DummyEngine) so the `test_xordemo`
pipeline KAT can be replayed.  The group elements are `Scalar<E>` (domain.rs:224),
i.e. plain field elements as canonical ints.
"""
from __future__ import annotations

from .multiexp import PolynomialDegreeTooLarge


def bitreverse(n: int, l: int) -> int:
    """domain.rs:274-281"""
    r = 0
    for _ in range(l):
        r = (r << 1) | (n & 1)
        n >>= 1
    return r


def serial_fft(F, a, omega, log_n):
    """domain.rs:272-315: in-place bit-reversal then log_n DIT stages. natural in, natural out."""
    n = len(a)
    assert n == 1 << log_n
    for k in range(n):
        rk = bitreverse(k, log_n)
        if k < rk:
            a[rk], a[k] = a[k], a[rk]
    m = 1
    p = F.p
    for _ in range(log_n):
        w_m = pow(omega, n // (2 * m), p)
        k = 0
        while k < n:
            w = 1
            for j in range(m):
                t = (a[k + j + m] * w) % p
                tmp = a[k + j] - t
                if tmp < 0:
                    tmp += p
                a[k + j + m] = tmp
                s = a[k + j] + t
                if s >= p:
                    s -= p
                a[k + j] = s
                w = (w * w_m) % p
            k += 2 * m
        m *= 2


def parallel_fft(F, a, omega, log_n, log_cpus):
    """domain.rs:317-374: 2^log_cpus twiddled sub-FFTs + transposing gather."""
    assert log_n >= log_cpus
    p = F.p
    num_cpus = 1 << log_cpus
    log_new_n = log_n - log_cpus
    tmp = [[0] * (1 << log_new_n) for _ in range(num_cpus)]
    new_omega = pow(omega, num_cpus, p)
    for j in range(num_cpus):
        t = tmp[j]
        omega_j = pow(omega, j, p)
        omega_step = pow(omega, j << log_new_n, p)
        elt = 1
        for i in range(1 << log_new_n):
            for s in range(num_cpus):
                idx = (i + (s << log_new_n)) % (1 << log_n)
                t[i] = (t[i] + a[idx] * elt) % p
                elt = (elt * omega_step) % p
            elt = (elt * omega_j) % p
        serial_fft(F, t, new_omega, log_new_n)
    mask = (1 << log_cpus) - 1
    for idx in range(len(a)):
        a[idx] = tmp[idx & mask][idx >> log_cpus]


def best_fft(F, a, omega, log_n, log_cpus=0):
    """domain.rs:261-270"""
    if log_n <= log_cpus:
        serial_fft(F, a, omega, log_n)
    else:
        parallel_fft(F, a, omega, log_n, log_cpus)


class EvaluationDomain:
    """domain.rs:26-189"""

    def __init__(self, F, coeffs, log_cpus=0):
        # from_coeffs, domain.rs:48-81
        self.F = F
        m = 1
        exp = 0
        while m < len(coeffs):
            m *= 2
            exp += 1
            if exp >= F.S:
                raise PolynomialDegreeTooLarge()
        omega = F.root_of_unity
        for _ in range(exp, F.S):
            omega = F.sqr(omega)
        self.coeffs = list(coeffs) + [0] * (m - len(coeffs))
        self.exp = exp
        self.omega = omega
        self.omegainv = F.inv(omega)
        self.geninv = F.inv(F.generator)
        self.minv = F.inv(m % F.p)
        self.log_cpus = log_cpus

    @classmethod
    def from_coeffs(cls, F, coeffs, log_cpus=0):
        return cls(F, coeffs, log_cpus)

    def into_coeffs(self):
        return self.coeffs

    def fft(self):
        best_fft(self.F, self.coeffs, self.omega, self.exp, min(self.log_cpus, self.exp))

    def ifft(self):
        best_fft(self.F, self.coeffs, self.omegainv, self.exp, min(self.log_cpus, self.exp))
        p = self.F.p
        minv = self.minv
        self.coeffs = [(v * minv) % p for v in self.coeffs]

    def distribute_powers(self, g):
        p = self.F.p
        u = 1
        out = []
        for v in self.coeffs:
            out.append((v * u) % p)
            u = (u * g) % p
        self.coeffs = out

    def coset_fft(self):
        self.distribute_powers(self.F.generator)
        self.fft()

    def icoset_fft(self):
        self.ifft()
        self.distribute_powers(self.geninv)

    def z(self, tau):
        return self.F.sub(self.F.pow(tau, len(self.coeffs)), 1)

    def divide_by_z_on_coset(self):
        i = self.F.inv(self.z(self.F.generator))
        p = self.F.p
        self.coeffs = [(v * i) % p for v in self.coeffs]

    def mul_assign(self, other):
        assert len(self.coeffs) == len(other.coeffs)
        p = self.F.p
        self.coeffs = [(a * b) % p for a, b in zip(self.coeffs, other.coeffs)]

    def sub_assign(self, other):
        assert len(self.coeffs) == len(other.coeffs)
        p = self.F.p
        self.coeffs = [(a - b) % p for a, b in zip(self.coeffs, other.coeffs)]


def h_coefficients(F, a, b, c):
    """The H-polynomial block of create_proof, groth16/prover.rs:256-287.

    a, b, c: evaluation vectors (canonical ints).  Returns the m-1 coefficients
    (canonical ints == into_repr) that feed the H multiexp.
    """
    da = EvaluationDomain(F, a)
    db = EvaluationDomain(F, b)
    dc = EvaluationDomain(F, c)
    da.ifft(); da.coset_fft()
    db.ifft(); db.coset_fft()
    dc.ifft(); dc.coset_fft()
    da.mul_assign(db)
    da.sub_assign(dc)
    da.divide_by_z_on_coset()
    da.icoset_fft()
    out = da.into_coeffs()
    return out[: len(out) - 1]
