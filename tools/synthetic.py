"""Seeded synthetic inputs of the shapes BASELINE.json names (no oracle and no product imports: plain numpy).

Shared by bench.py and the parity tests so that both sides of a comparison consume the same arrays."""
import numpy as np

FR_MODULUS = 0x73eda753299d7d483339d80809a1d80553bda402fffe5bfeffffffff00000001

# Sapling Spend: 98 777 + 8 constraints -> m = 2^17; 8 inputs (ONE + 7 public), 98 638 aux variables;
# A query 8 + 85 382 bases, B queries 1 + 61 299 bases (prover.rs:425-792 dump ranges; circuit/sapling/mod.rs:715)
SPEND_SHAPE = dict(n_con=98785, n_in=8, n_aux=98638, a_dense=85382, b_in_dense=1, b_aux_dense=61299)
# Sprout JoinSplit on Groth16: 1 989 085 constraints -> m = 2^21 (sapling-crypto/src/circuit/sprout/mod.rs:465); the variable
# counts and densities are not published: taken proportional to Spend's
SPROUT_SHAPE = dict(n_con=1989085, n_in=10, n_aux=1986000, a_dense=1719000, b_in_dense=1, b_aux_dense=1234000)


def random_scalars(rng, n):
    """uniform canonical Fr scalars (n, 4) uint64 -- 255-bit rejection sampling like fr.rs:255-268"""
    mod = [(FR_MODULUS >> (64 * i)) & 0xFFFFFFFFFFFFFFFF for i in range(4)]
    out = rng.integers(0, 1 << 64, size=(n, 4), dtype=np.uint64)
    out[:, 3] &= np.uint64((1 << 63) - 1)
    while True:
        ge = np.zeros(n, dtype=bool)
        eq = np.ones(n, dtype=bool)
        for l in (3, 2, 1, 0):
            ge |= eq & (out[:, l] > np.uint64(mod[l]))
            eq &= out[:, l] == np.uint64(mod[l])
        bad = ge | eq
        k = int(bad.sum())
        if k == 0:
            return out
        fresh = rng.integers(0, 1 << 64, size=(k, 4), dtype=np.uint64)
        fresh[:, 3] &= np.uint64((1 << 63) - 1)
        out[bad] = fresh


def witness_scalars(rng, n):
    """witness-like assignment: half of the values are 0 / 1 (boolean wires dominate a Sapling witness), the rest uniform"""
    v = random_scalars(rng, n)
    small = rng.random(n) < 0.5
    v[small] = 0
    v[small, 0] = rng.integers(0, 2, size=int(small.sum()), dtype=np.uint64)
    return v


def density(rng, n, total):
    d = np.zeros(n, dtype=np.uint8)
    d[rng.choice(n, size=total, replace=False)] = 1
    return d


def base_multipliers(rng, n):
    """k_i (64-bit, non-zero) as (n, 4) canonical limbs: synthetic bases are [k_i] * generator"""
    k = np.zeros((n, 4), dtype=np.uint64)
    k[:, 0] = rng.integers(1, 1 << 64, size=n, dtype=np.uint64)
    return k


def spend_assignment(rng, shape=SPEND_SHAPE):
    """One synthesized ProvingAssignment of the given shape (prover.rs:84-190): a, b, c evaluation vectors (Montgomery residues),
    input / aux assignments (canonical, inputs[0] = ONE), the three density maps."""
    s = shape
    a, b, c = (random_scalars(rng, s["n_con"]) for _ in range(3))
    inputs, aux = witness_scalars(rng, s["n_in"]), witness_scalars(rng, s["n_aux"])
    inputs[0] = (1, 0, 0, 0)
    return dict(a=a, b=b, c=c, inputs=inputs, aux=aux, a_aux_density=density(rng, s["n_aux"], s["a_dense"]),
                b_input_density=density(rng, s["n_in"], s["b_in_dense"]), b_aux_density=density(rng, s["n_aux"], s["b_aux_dense"]))


def crs_sizes(shape=SPEND_SHAPE):
    """lengths of the h, l, a, b_g1, b_g2 query vectors (groth16/mod.rs:215-238) for a circuit of this shape"""
    s = shape
    m = 1 << (s["n_con"] - 1).bit_length()
    return dict(h=m - 1, l=s["n_aux"], a=s["n_in"] + s["a_dense"], b_g1=s["b_in_dense"] + s["b_aux_dense"], b_g2=s["b_in_dense"] + s["b_aux_dense"])
