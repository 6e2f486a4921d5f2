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


def chunk_grid(n_chunks: int) -> tuple[int, int, int]:
    """n_chunks = kx * ky * kz, as cubic as the factorisation allows."""
    best = (n_chunks, 1, 1)
    for kx in range(1, n_chunks + 1):
        if n_chunks % kx:
            continue
        for ky in range(1, n_chunks // kx + 1):
            if (n_chunks // kx) % ky:
                continue
            kz = n_chunks // kx // ky
            if max(kx, ky, kz) < max(best):
                best = (kx, ky, kz)
    return best


def split_slabs(s: Snapshot, n_chunks: int, box: float = 100000.0) -> list[Snapshot]:
    """The chunk files of a full-box AHF run: one per spatial sub-volume (compact blocks of a kx x ky x kz decomposition, the shape
    MPI-AHF's domains have; the reference's buffer-halo exchange ships the shell around each task's volume)."""
    if s.pos is None:                                         # catalogue without positions (all halos at the box centre when written): round robin
        slab = np.arange(s.n_halos, dtype=np.int64) % n_chunks
    else:
        kx, ky, kz = chunk_grid(n_chunks)
        cell = [np.minimum((s.pos[:, a] / box * k).astype(np.int64), k - 1) for a, k in enumerate((kx, ky, kz))]
        slab = (cell[0] * ky + cell[1]) * kz + cell[2]
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


def synth_chain(seed: int, n_snaps: int, n_halos: int, n_tuples: int, multi_species: bool = False, p_vanish: float = 0.03,
                p_merge: float = 0.02, box: float = 100000.0) -> list[Snapshot]:
    """A chain of n_snaps DISTINCT snapshots (ascending snapshot number = forward in time) of one evolving halo population, vectorised
    (BASELINE.json configs[3] and [4]): per step every particle stays in its halo (88 %), moves to a neighbouring
    halo (6 %) or leaves and is replaced by a newly accreted ID (6 %); 2 % of the halos merge into another one; p_vanish of them are
    missing from the catalogues for 1-4 snapshots and come back (what the token / orphan rule tracks); 15 % are subhalos, listed under
    their host too; the halo labels are permuted at every snapshot.  multi_species: types {0 gas, 1 DM, 4 star} (25 / 65 / 10 %), 2 % of
    the gas particles of a halo turn into stars per step (same ID, new type)."""
    rng = np.random.default_rng(seed)
    H = n_halos
    s = np.maximum(20.0, 20.0 * rng.random(H) ** (-1.0 / 0.9))
    s = np.minimum(s, 1e-2 * n_tuples)
    s = np.maximum(20, (s * (n_tuples / 1.15 / s.sum())).astype(np.int64))
    P = int(s.sum())
    lineage = np.repeat(np.arange(H, dtype=np.int64), s)            # halo (by lineage number) of every particle slot
    next_id = np.uint64(1)
    # IDs are scattered over about twice the bound count (half of the particles of the box are bound), new IDs come from the same range
    span = np.uint64(2 * P + 2 * P // 10 * n_snaps + 64)
    mult = np.uint64(0x9E3779B97F4A7C15 | 1)

    def fresh(n):
        nonlocal next_id
        ctr = np.arange(int(next_id), int(next_id) + n, dtype=np.uint64)
        next_id = np.uint64(int(next_id) + n)
        return ctr                                                  # relabelled below by a fixed permutation of [1, span]
    pid = fresh(P)
    ptype = (rng.choice(np.array([0, 1, 4], np.uint8), size=P, p=[0.25, 0.65, 0.10]) if multi_species else np.ones(P, np.uint8))
    is_sub = rng.random(H) < 0.15
    is_sub[0] = False
    host = np.arange(H, dtype=np.int64)
    last = np.maximum.accumulate(np.where(~is_sub, np.arange(H), 0))
    host[is_sub] = last[is_sub]
    alive = np.ones(H, bool)                                        # merged-away lineages are dead for good
    hidden_until = np.zeros(H, np.int64)                            # lineage is missing from the catalogues of steps < hidden_until
    pos = rng.random((H, 3)) * box
    perm_ids = None
    snaps = []
    for step in range(n_snaps):
        if step > 0:
            u = rng.random(P)
            moved = (u >= 0.88) & (u < 0.94)
            d = rng.integers(1, 4, size=P) * rng.choice(np.array([-1, 1]), size=P)
            tgt = np.clip(lineage + d, 0, H - 1)
            ok = moved & alive[tgt]
            lineage = np.where(ok, tgt, lineage)
            gone = u >= 0.94
            pid[gone] = fresh(int(gone.sum()))
            if multi_species:
                ptype[gone] = rng.choice(np.array([0, 1, 4], np.uint8), size=int(gone.sum()), p=[0.25, 0.65, 0.10])
                gas = (ptype == 0) & ~gone
                ptype[gas & (rng.random(P) < 0.02)] = 4
            # mergers: a lineage is absorbed by another live, non-sub lineage
            is_host = np.bincount(host[is_sub & alive], minlength=H) > 0
            cand = np.nonzero(alive & ~is_sub & ~is_host & (rng.random(H) < p_merge))[0]
            into = rng.integers(H, size=len(cand))
            good = (into != cand) & alive[into] & ~is_sub[into] & ~np.isin(into, cand)
            merge_to = np.arange(H, dtype=np.int64)
            merge_to[cand[good]] = into[good]
            lineage = merge_to[lineage]
            alive[cand[good]] = False
            # halos that drop out of the catalogue for a few snapshots
            drop = alive & (hidden_until <= step) & (rng.random(H) < p_vanish)
            hidden_until[drop] = step + rng.integers(1, 5, size=int(drop.sum()))
            pos = np.mod(pos + rng.standard_normal((H, 3)) * 50.0, box)
        visible = alive & (hidden_until <= step)
        vis_idx = np.nonzero(visible)[0]
        label = np.full(H, -1, np.int64)
        label[vis_idx] = rng.permutation(len(vis_idx))
        own = visible[lineage]
        p_pid, p_lab, p_typ = pid[own], label[lineage[own]], ptype[own]
        # double membership under a visible host
        sub = is_sub[lineage] & own & visible[host[lineage]]
        p_pid = np.concatenate([p_pid, pid[sub]])
        p_lab = np.concatenate([p_lab, label[host[lineage[sub]]]])
        p_typ = np.concatenate([p_typ, ptype[sub]])
        order = np.argsort(p_lab, kind="stable")
        n_h = len(vis_idx)
        cnt = np.bincount(p_lab, minlength=n_h)
        off = np.zeros(n_h + 1, np.int64)
        np.cumsum(cnt, out=off[1:])
        number = 100 + step
        ids = np.uint64(number) * np.uint64(10**12) + np.arange(1, n_h + 1, dtype=np.uint64)
        hpos = np.empty((n_h, 3), np.float32)
        hpos[label[vis_idx]] = pos[vis_idx]
        # counter -> particle ID: odd multiplier mod 2^k is a bijection; k = bits of the ID range
        k = int(span).bit_length()
        scat = ((p_pid[order] * mult) & np.uint64((1 << k) - 1)) + np.uint64(1)
        snaps.append(Snapshot(number, round(0.05 * (n_snaps - 1 - step), 3), ids, cnt.astype(np.int32), off, scat, p_typ[order], hpos,
                              np.zeros((n_h, 3), np.float32)))
    return snaps
