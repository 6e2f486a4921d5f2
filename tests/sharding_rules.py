"""Host-side rules of the multi-GPU layout: a numpy mirror of metrocpp_b200/csrc/comm.cu for tests/test_sharding_gloo.py (test
infrastructure - the product routes on the device; tests/test_multi_gpu.py checks comm.cu itself against the oracle).

* tuples of BOTH catalogues are sharded by particle-ID range, so a particle's memberships all
  live on one rank and sort + join need no communication;
* a reduced partial pair count (A, B, types, n) is sent to owner(A) (forward trees) and owner(B)
  (backward trees); owners are contiguous halo-index ranges of ceil(H / world) halos.
"""
from __future__ import annotations

import numpy as np


def range_chunk(n_halos: int, world: int) -> int:
    return (n_halos + world - 1) // world if n_halos > 0 else 1


def owner_of_halo(halo: np.ndarray, n_halos: int, world: int) -> np.ndarray:
    return (np.asarray(halo, np.int64) // range_chunk(n_halos, world)).astype(np.int64)


def owned_range(n_halos: int, world: int, rank: int) -> tuple[int, int]:
    c = range_chunk(n_halos, world)
    return min(n_halos, c * rank), min(n_halos, c * (rank + 1))


def pid_shard_bounds(pid_bits: int, world: int, rank: int) -> tuple[int, int]:
    """[lo, hi) of the particle-ID range of `rank` (world must be a power of two)."""
    assert world & (world - 1) == 0, "world must be a power of two"
    per = (1 << pid_bits) // world
    return per * rank, per * (rank + 1)


def pid_shard_mask(pid: np.ndarray, pid_bits: int, world: int, rank: int) -> np.ndarray:
    lo, hi = pid_shard_bounds(pid_bits, world, rank)
    return (pid >= np.uint64(lo)) & (pid < np.uint64(hi))


def route_records(rec_a: np.ndarray, rec_b: np.ndarray, counts: np.ndarray, n_halos0: int, n_halos1: int, world: int):
    """Split partial pair records into per-destination lists: ([by owner(A)], [by owner(B)])."""
    oa, ob = owner_of_halo(rec_a, n_halos0, world), owner_of_halo(rec_b, n_halos1, world)
    by_a = [(rec_a[oa == r], rec_b[oa == r], counts[oa == r]) for r in range(world)]
    by_b = [(rec_a[ob == r], rec_b[ob == r], counts[ob == r]) for r in range(world)]
    return by_a, by_b
