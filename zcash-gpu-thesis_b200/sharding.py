"""Multi-GPU partitioning of the prover hot path (one process per GPU).

MSM shards by contiguous base range: rank r of g owns (base, exponent) pairs [lo, hi) of the n pairs, runs the
full single-GPU multiexp on them and contributes one Jacobian partial (144 B for G1, 288 B for G2); the partials
are all-gathered (NCCL inside libb200zk.so on the GPU path) and summed in rank order.  Independent NTTs and whole
proofs go round-robin.  Only plain index arithmetic lives here so that it can be tested on CPU (gloo)."""
from __future__ import annotations


def shard_range(n: int, rank: int, world: int):
    """Contiguous, balanced split of n pairs: the first n % world ranks get one extra pair."""
    assert 0 <= rank < world
    base, extra = divmod(n, world)
    lo = rank * base + min(rank, extra)
    hi = lo + base + (1 if rank < extra else 0)
    return lo, hi


def shard_density(density, lo: int, hi: int, base_offset: int = 0):
    """For a density-mapped multiexp the base cursor of a shard starts after the set bits that precede it
    (multiexp.rs:174-196 consumes one base per set bit).  Returns (density[lo:hi], base_offset + popcount(density[:lo]))."""
    if density is None:
        return None, base_offset + lo
    before = int(sum(1 for b in density[:lo] if b))
    return density[lo:hi], base_offset + before


def round_robin(items: int, rank: int, world: int):
    """Indices of the independent work items (the a/b/c NTTs of a proof, or whole proofs of a batch) for this rank."""
    return list(range(rank, items, world))


def lockstep_groups(items: int, lockstep: int):
    """The batch cut into consecutive groups of at most `lockstep` proofs (one b200zk_groth16_prove_batch call each)."""
    lockstep = max(1, lockstep)
    return [list(range(lo, min(lo + lockstep, items))) for lo in range(0, items, lockstep)]


def prove_on_devices(workers, params_per_worker, assignments, lockstep: int = 8):
    """A batch of proofs over several GPUs driven by ONE host process (what an FFI caller of the reference would be): the
    CRS is replicated (`params_per_worker[i]` lives on `workers[i]`'s device), the lock-step groups of the batch go round-robin
    over the contexts, one host thread per context.  Returns the proofs in the order of `assignments`."""
    import threading

    from .bellman import create_proofs_from_assignments

    groups = lockstep_groups(len(assignments), lockstep)
    out, errs = [None] * len(assignments), []

    def run(i):
        try:
            for g in round_robin(len(groups), i, len(workers)):
                idx = groups[g]
                for j, p in zip(idx, create_proofs_from_assignments(workers[i], params_per_worker[i], [assignments[k] for k in idx], lockstep)):
                    out[j] = p
        except Exception as e:  # noqa: BLE001
            errs.append(e)

    threads = [threading.Thread(target=run, args=(i,)) for i in range(len(workers))]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    if errs:
        raise errs[0]
    return out
