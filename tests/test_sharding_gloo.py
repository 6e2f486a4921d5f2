"""N>1 host logic on CPU (gloo, world_size 2): pid-range sharding + owner routing + summing the
partial pair counts reproduces the unsharded pair counts (what the NCCL path does with device
buffers, csrc/comm.cu)."""
import os

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

import metrocpp_b200 as mb
from tests import sharding_rules as sharding


def pair_counts(c0, c1):
    """numpy hash-free join: distinct (A, B) -> number of shared tuples (cross product per pid)."""
    o0, o1 = np.argsort(c0.pid, kind="stable"), np.argsort(c1.pid, kind="stable")
    p0, h0, p1, h1 = c0.pid[o0], c0.halo[o0], c1.pid[o1], c1.halo[o1]
    lo, hi = np.searchsorted(p1, p0, "left"), np.searchsorted(p1, p0, "right")
    reps = hi - lo
    a = np.repeat(h0, reps)
    idx = np.concatenate([np.arange(l, h) for l, h in zip(lo[reps > 0], hi[reps > 0])]) if reps.sum() else np.zeros(0, np.int64)
    b = h1[idx]
    key = a.astype(np.uint64) << np.uint64(32) | b.astype(np.uint64)
    k, n = np.unique(key, return_counts=True)
    return (k >> np.uint64(32)).astype(np.int64), (k & np.uint64(0xFFFFFFFF)).astype(np.int64), n.astype(np.int64)


def worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    prm = mb.make_params(nptypes=2)
    bits = 18
    spec = mb.SynthSpec(seed=5, n_halos=300, n_tuples=40_000, pid_bits=bits, shard=0, n_shards=1)
    c0, c1 = mb.synth_pair_host(prm, spec, 0), mb.synth_pair_host(prm, spec, 1)
    # my shard: must equal what the generator emits for (shard=rank, n_shards=world)
    m0, m1 = sharding.pid_shard_mask(c0.pid, bits, world, rank), sharding.pid_shard_mask(c1.pid, bits, world, rank)
    mine = mb.SynthSpec(seed=5, n_halos=300, n_tuples=40_000, pid_bits=bits, shard=rank, n_shards=world)
    g0 = mb.synth_pair_host(prm, mine, 0)
    assert np.array_equal(np.sort(g0.pid), np.sort(c0.pid[m0]))
    s0 = mb.HostCatalogue(c0.halo_id, c0.npart, c0.n_orphan_steps, c0.is_token, c0.pid[m0], c0.halo[m0], None)
    s1 = mb.HostCatalogue(c1.halo_id, c1.npart, c1.n_orphan_steps, c1.is_token, c1.pid[m1], c1.halo[m1], None)
    a, b, n = pair_counts(s0, s1)                                   # local partial counts
    by_a, by_b = sharding.route_records(a, b, n, c0.n_halos, c1.n_halos, world)
    recv_a, recv_b = [None] * world, [None] * world                  # the all-to-all
    for src in range(world):
        box = [by_a if rank == src else None, by_b if rank == src else None]
        dist.broadcast_object_list(box, src=src)
        recv_a[src], recv_b[src] = box[0][rank], box[1][rank]
    fa, fb, fn = pair_counts(c0, c1)                                 # the unsharded truth
    for recv, owner_of, H in ((recv_a, fa, c0.n_halos), (recv_b, fb, c1.n_halos)):
        lo, hi = sharding.owned_range(H, world, rank)
        ka = np.concatenate([r[0] for r in recv]); kb = np.concatenate([r[1] for r in recv]); kn = np.concatenate([r[2] for r in recv])
        key = ka.astype(np.uint64) << np.uint64(32) | kb.astype(np.uint64)
        uk, inv = np.unique(key, return_inverse=True)
        summed = np.bincount(inv, weights=kn, minlength=uk.shape[0]).astype(np.int64)
        sel = (owner_of >= lo) & (owner_of < hi)
        tk = fa[sel].astype(np.uint64) << np.uint64(32) | fb[sel].astype(np.uint64)
        assert np.array_equal(uk, tk) and np.array_equal(summed, fn[sel])
    out.put((rank, int(n.sum())))
    dist.destroy_process_group()


def test_two_rank_sharded_pair_counts_equal_unsharded():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 400
    procs = [ctx.Process(target=worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(180)
        assert p.exitcode == 0
    got = sorted(q.get(timeout=5) for _ in range(2))
    assert [g[0] for g in got] == [0, 1] and all(g[1] > 0 for g in got)


@pytest.mark.parametrize("H,world", [(10, 4), (1, 2), (0, 2), (1000003, 8)])
def test_owner_ranges_partition_the_halos(H, world):
    covered = []
    for r in range(world):
        lo, hi = sharding.owned_range(H, world, r)
        assert 0 <= lo <= hi <= H
        covered += list(range(lo, min(hi, lo + 3))) + list(range(max(lo, hi - 3), hi))
        if hi > lo:
            assert np.all(sharding.owner_of_halo(np.array([lo, hi - 1]), H, world) == r)
    assert sum(sharding.owned_range(H, world, r)[1] - sharding.owned_range(H, world, r)[0] for r in range(world)) == H
