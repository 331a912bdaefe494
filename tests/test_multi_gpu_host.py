"""world_size-2 `gloo` test of the multi-GPU host logic: base-range sharding, density-aware base cursors, gather of
the per-shard Jacobian partials and their sum == the unsharded multiexp.  The shard MSMs are computed by the CPU
oracle here (no GPU in this container); the GPU path uses the same shard_range / shard_density and the same
gather-then-add order inside b200zk_allgather_sum_dev."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_path):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import zcash_gpu_thesis_b200  # noqa: F401  (package import must work without a GPU)
    from zcash_gpu_thesis_b200.sharding import shard_density, shard_range
    from oracle import cref
    from tests import util

    r = util.rng(77)  # same inputs on every rank
    n = 1000
    xy, ks = util.random_bases("g1", r, 700)
    exps = util.random_fr_repr(r, n)
    density = (r.random(n) < 0.6).astype(np.uint8)
    density[np.nonzero(density)[0][650:]] = 0
    lo, hi = shard_range(n, rank, world)
    d_shard, off = shard_density(density, lo, hi, base_offset=20)
    st, partial = cref.multiexp("g1", xy, exps[lo:hi], density=d_shard, base_offset=off)
    assert st == 0
    t = torch.from_numpy(partial.view(np.int64).copy())
    gathered = [torch.zeros_like(t) for _ in range(world)]
    dist.all_gather(gathered, t)
    acc = gathered[0].numpy().view(np.uint64)
    for g in gathered[1:]:
        acc = cref.point_op("g1", "add", acc, g.numpy().view(np.uint64))
    if rank == 0:
        st, full = cref.multiexp("g1", xy, exps, density=density, base_offset=20)
        a1, i1 = cref.into_affine("g1", acc)
        a2, i2 = cref.into_affine("g1", full)
        ok = st == 0 and i1 == i2 and np.array_equal(a1, a2)
        open(out_path, "w").write("ok" if ok else "mismatch")
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_multiexp_world2_gloo(tmp_path):
    out = tmp_path / "result.txt"
    mp.spawn(_worker, args=(2, _free_port(), str(out)), nprocs=2, join=True)
    assert out.read_text() == "ok"


def test_shard_ranges_partition():
    from zcash_gpu_thesis_b200.sharding import round_robin, shard_range

    for n in (0, 1, 7, 8, 1000, (1 << 24) + 3):
        for world in (1, 2, 4, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1
    assert sorted(sum((round_robin(7, r, 3) for r in range(3)), [])) == list(range(7))


def test_multi_plan_matches_shard_density():
    """b200zk_multi_plan (the host logic of the one-process multi-GPU multiexp, callable without a GPU): the exponent cuts and
    per-shard base cursors give, shard by shard, exactly the (base, exponent) pairs of the unsharded multiexp -- checked by
    replaying the Source semantics (multiexp.rs:42-68, 174-196) in plain Python."""
    import zcash_gpu_thesis_b200 as zk

    rng = np.random.default_rng(5)

    def pairs_unsharded(n_bases, off, density, n_exp):
        out, idx = [], off
        for i in range(n_exp):
            if density is not None and not density[i]:
                continue
            out.append((i, idx if idx < n_bases else "EOF"))
            idx += 1
        return out

    for trial in range(200):
        nd = int(rng.integers(1, 6))
        n_bases = int(rng.integers(0, 60))
        base = n_bases // nd
        bounds = [d * base + min(d, n_bases % nd) for d in range(nd + 1)]
        n_exp = int(rng.integers(0, 80))
        off = int(rng.integers(0, n_bases + 3))
        density = None if trial % 3 == 0 else (rng.random(n_exp) < rng.random()).astype(np.uint8)
        e_lo, loc = zk.multi_plan(bounds, off, density, n_exp)
        assert e_lo[0] == 0 and e_lo[-1] == n_exp and all(e_lo[d] <= e_lo[d + 1] for d in range(nd))
        got = []
        for d in range(nd):
            n_shard = bounds[d + 1] - bounds[d]
            idx = loc[d]
            for i in range(e_lo[d], e_lo[d + 1]):
                if density is not None and not density[i]:
                    continue
                got.append((i, bounds[d] + idx if idx < n_shard else "EOF"))
                idx += 1
        want = pairs_unsharded(n_bases, off, density, n_exp)
        assert got == want, (bounds, off, n_exp, None if density is None else density.tolist())
