"""Synthetic full-box snapshot pair in plain numpy.  TEST INFRASTRUCTURE ONLY (inputs of the CPU reference arm).

Same population model as the GPU workload's generator (power-law halo sizes, 15 % subhalos whose particles are
also listed under their host, per particle 88 % stay / 6 % move to a neighbouring halo / 6 % are replaced by fresh IDs, halo labels
permuted between the snapshots), written independently of the product library so that the reference arm of bench.py never touches
libmetro_b200.so.  Halo positions are uniform in the box and a progenitor sits where its descendant sits (+ a small drift), which is
what the reference's grid / buffer-halo exchange between MPI tasks assumes (reference src/Grid.cpp:258-299).
"""
from __future__ import annotations

import numpy as np

from .ahf_io import Snapshot


def synth_pair(seed: int, n_halos: int, n_tuples: int, box: float = 100000.0) -> list[Snapshot]:
    """-> [earlier (number 1, z = 0.1), later (number 2, z = 0)]"""
    rng = np.random.default_rng(seed)
    H = n_halos
    # halo sizes: s = max(20, 20 u^(-1/0.9)), capped, rescaled to the tuple total
    s = np.maximum(20.0, 20.0 * rng.random(H) ** (-1.0 / 0.9))
    s = np.minimum(s, 1e-3 * n_tuples * 10)
    s = np.maximum(20, (s * (n_tuples / 1.15 / s.sum())).astype(np.int64))
    off = np.zeros(H + 1, np.int64)
    np.cumsum(s, out=off[1:])
    P = int(off[-1])
    pid_pool = rng.permutation(np.arange(1, 2 * P + 1, dtype=np.uint64))      # about half of all IDs are bound
    own_pid = pid_pool[:P]
    own_halo = np.repeat(np.arange(H, dtype=np.int64), s)
    fresh = pid_pool[P:]
    # subhalos: 15 % of the halos (never halo 0) are subhalos of the nearest preceding non-subhalo
    is_sub = rng.random(H) < 0.15
    is_sub[0] = False
    host = np.arange(H, dtype=np.int64)
    last = np.maximum.accumulate(np.where(~is_sub, np.arange(H), 0))
    host[is_sub] = last[is_sub]

    def catalogue(pid, halo, n_h, number, z, ids_base, pos):
        # AHF double membership: a subhalo's particles are listed under the host too
        sub = is_sub_of[halo] != halo if n_h == len(is_sub_of) else np.zeros(len(halo), bool)
        pid_all = np.concatenate([pid, pid[sub]])
        halo_all = np.concatenate([halo, is_sub_of[halo[sub]]]) if sub.any() else halo
        order = np.argsort(halo_all, kind="stable")
        pid_all, halo_all = pid_all[order], halo_all[order]
        cnt = np.bincount(halo_all, minlength=n_h)
        o = np.zeros(n_h + 1, np.int64)
        np.cumsum(cnt, out=o[1:])
        ids = np.uint64(ids_base) + np.arange(1, n_h + 1, dtype=np.uint64)
        return Snapshot(number, z, ids, cnt.astype(np.int32), o, pid_all, np.ones(len(pid_all), np.uint8), pos.astype(np.float32),
                        (rng.standard_normal((n_h, 3)) * 200.0).astype(np.float32))

    pos_later = rng.random((H, 3)) * box
    is_sub_of = host
    later = catalogue(own_pid, own_halo, H, 2, 0.0, 2 * 10**12, pos_later)
    # earlier snapshot: per particle 88 % same halo, 6 % moved to halo +-(1..3), 6 % replaced by a fresh (never bound later) ID
    u = rng.random(P)
    d = rng.integers(1, 4, size=P) * rng.choice(np.array([-1, 1]), size=P)
    src = np.where((u >= 0.88) & (u < 0.94), np.clip(own_halo + d, 0, H - 1), own_halo)
    pid1 = own_pid.copy()
    gone = u >= 0.94
    pid1[gone] = fresh[: int(gone.sum())]
    perm = rng.permutation(H)                                   # earlier label of the later halo h
    pos_earlier = np.empty_like(pos_later)
    pos_earlier[perm] = np.mod(pos_later + rng.standard_normal((H, 3)) * 50.0, box)
    sub1 = np.arange(H, dtype=np.int64)
    sub1[perm] = perm[host]                                     # the same host / subhalo relation under the earlier labels
    is_sub_of = sub1
    earlier = catalogue(pid1, perm[src], H, 1, 0.1, 1 * 10**12, pos_earlier)
    return [earlier, later]


def split_slabs(s: Snapshot, n_chunks: int, box: float = 100000.0) -> list[Snapshot]:
    """The chunk files of a full-box AHF run: one per spatial sub-volume (x slabs)."""
    if s.pos is None:                                         # catalogue without positions (all halos at the box centre when written): round robin
        slab = np.arange(s.n_halos, dtype=np.int64) % n_chunks
    else:
        slab = np.minimum((s.pos[:, 0] / box * n_chunks).astype(np.int64), n_chunks - 1)
    tuple_halo = s.halo_index_per_tuple()
    out = []
    for c in range(n_chunks):
        idx = np.nonzero(slab == c)[0]
        keep = slab[tuple_halo] == c
        cnt = np.diff(s.offsets)[idx]
        o = np.zeros(len(idx) + 1, np.int64)
        np.cumsum(cnt, out=o[1:])
        out.append(Snapshot(s.number, s.redshift, s.halo_id[idx], s.npart[idx], o, s.pid[keep], s.ptype[keep],
                            s.pos[idx] if s.pos is not None else None, s.vel[idx] if s.vel is not None else None))
    return out
