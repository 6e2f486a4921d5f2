"""Small seeded AHF snapshot chains for the parity tests (host-only helpers).

Forward-time toy model of a halo population: halos keep ~90 % of their particles per
step, accrete fresh ones, merge, become subhalos (their particles are then ALSO listed
under the host, AHF-style double membership), vanish for a few snapshots and return
(exercises the token/orphan rule of reference src/MergerTree.cpp:575-617) and, in the
multi-species variant, gas particles (type 0) turn into stars (type 4).
"""
from __future__ import annotations

import numpy as np

from oracle.ahf_io import Snapshot


def _fresh(rng, state, n):
    lo = state["next_pid"]
    state["next_pid"] = lo + n
    # scatter the consecutive counter over a 2^40 range with an odd multiplier (bijective mod 2^40)
    x = (np.arange(lo, lo + n, dtype=np.uint64) * np.uint64(0x9E3779B97F) + np.uint64(12345)) & np.uint64((1 << 40) - 1)
    return x + np.uint64(1)


def make_chain(seed: int = 1, n_snaps: int = 4, n_halos: int = 60, mean_size: int = 300,
               multi_species: bool = False, p_merge: float = 0.05, p_sub: float = 0.15,
               p_vanish: float = 0.06, first_number: int = 10, id_mode: str = "snap",
               min_size: int = 20) -> list[Snapshot]:
    """Returns snapshots in ascending snapshot number (= forward time)."""
    rng = np.random.default_rng(seed)
    state = {"next_pid": 1}
    halos = []      # dicts: parts (uint64), types (uint8), host (index in list or None)
    for _ in range(n_halos):
        size = int(max(min_size, min(20 * rng.random() ** (-1 / 0.9), 40 * mean_size)))
        parts = _fresh(rng, state, size)
        halos.append({"parts": parts, "types": _types(rng, size, multi_species), "sub_of": None})
    limbo = []      # (return_step, halo)
    snaps = []
    for step in range(n_snaps):
        number = first_number + step
        if step > 0:
            halos, limbo = _evolve(rng, state, halos, limbo, step, multi_species, p_merge, p_sub,
                                   p_vanish, mean_size, min_size)
        snaps.append(_catalogue(rng, halos, number, n_snaps - 1 - step, id_mode))
    return snaps


def _types(rng, n, multi):
    if not multi:
        return np.ones(n, np.uint8)
    return rng.choice(np.array([0, 1, 4], np.uint8), size=n, p=[0.25, 0.65, 0.10])


def _evolve(rng, state, halos, limbo, step, multi, p_merge, p_sub, p_vanish, mean_size, min_size):
    out = []
    n = len(halos)
    # returning halos
    still = []
    for (when, h) in limbo:
        if when <= step:
            out.append(h)
        else:
            still.append((when, h))
    limbo = still
    absorbed = {}
    order = rng.permutation(n)
    for i in order:
        h = halos[i]
        keep = rng.random(h["parts"].shape[0]) < 0.9
        parts = h["parts"][keep]
        types = h["types"][keep].copy()
        if multi:
            gas = np.nonzero(types == 0)[0]
            flip = gas[rng.random(gas.shape[0]) < 0.02]
            types[flip] = 4
        n_new = max(1, int(0.1 * h["parts"].shape[0]))
        parts = np.concatenate([parts, _fresh(rng, state, n_new)])
        types = np.concatenate([types, _types(rng, n_new, multi)])
        nh = {"parts": parts, "types": types, "sub_of": None}
        u = rng.random()
        if u < p_merge and out:
            tgt = out[int(rng.integers(len(out)))]
            tgt["parts"] = np.concatenate([tgt["parts"], parts])
            tgt["types"] = np.concatenate([tgt["types"], types])
            continue
        if u < p_merge + p_vanish:
            limbo.append((step + int(rng.integers(1, 5)), nh))
            continue
        if u < p_merge + p_vanish + p_sub and out:
            # become a subhalo of an existing (non-sub) halo
            cands = [k for k, o in enumerate(out) if o["sub_of"] is None]
            if cands:
                nh["sub_of"] = int(cands[int(rng.integers(len(cands)))])
        out.append(nh)
    # a few newborn halos
    for _ in range(max(1, n // 20)):
        size = int(max(min_size, min(20 * rng.random() ** (-1 / 0.9), 40 * mean_size)))
        out.append({"parts": _fresh(rng, state, size), "types": _types(rng, size, multi), "sub_of": None})
    return out, limbo


def _catalogue(rng, halos, number, z_index, id_mode):
    H = len(halos)
    lists_p, lists_t = [], []
    subs = {}                       # host index -> its subhalos, in list order
    for s in halos:
        if s["sub_of"] is not None:
            subs.setdefault(s["sub_of"], []).append(s)
    for k, h in enumerate(halos):
        p, t = [h["parts"]], [h["types"]]
        for s in subs.get(k, ()):
            p.append(s["parts"])
            t.append(s["types"])
        lists_p.append(np.concatenate(p))
        lists_t.append(np.concatenate(t))
    perm = rng.permutation(H)
    # remap sub_of through the permutation is not needed: membership is already materialised
    lists_p = [lists_p[i] for i in perm]
    lists_t = [lists_t[i] for i in perm]
    counts = np.array([x.shape[0] for x in lists_p], np.int64)
    offsets = np.zeros(H + 1, np.int64)
    np.cumsum(counts, out=offsets[1:])
    if id_mode == "snap":
        ids = np.uint64(number) * np.uint64(10**12) + np.arange(1, H + 1, dtype=np.uint64)
    elif id_mode == "hashed":       # newer AHF: hashed 64-bit IDs, not ordered like the file
        ids = rng.integers(1 << 40, 1 << 62, size=H, dtype=np.uint64)
        ids = np.unique(ids)
        while ids.shape[0] < H:
            ids = np.unique(np.concatenate([ids, rng.integers(1 << 40, 1 << 62, size=H, dtype=np.uint64)]))
        ids = rng.permutation(ids[:H])
    else:
        raise ValueError(id_mode)
    pos = (rng.random((H, 3)) * 100000.0).astype(np.float32)
    vel = (rng.standard_normal((H, 3)) * 200.0).astype(np.float32)
    return Snapshot(number=number, redshift=0.1 * z_index, halo_id=ids,
                    npart=counts.astype(np.int32), offsets=offsets,
                    pid=np.concatenate(lists_p) if H else np.zeros(0, np.uint64),
                    ptype=np.concatenate(lists_t) if H else np.zeros(0, np.uint8),
                    pos=pos, vel=vel)


def snapshot_from_lists(number, redshift, halo_ids, part_lists, type_lists=None, npart=None):
    """Hand-crafted snapshot from explicit per-halo particle lists."""
    H = len(halo_ids)
    counts = np.array([len(x) for x in part_lists], np.int64)
    offsets = np.zeros(H + 1, np.int64)
    np.cumsum(counts, out=offsets[1:])
    pid = np.concatenate([np.asarray(x, np.uint64) for x in part_lists]) if H else np.zeros(0, np.uint64)
    if type_lists is None:
        pt = np.ones(pid.shape[0], np.uint8)
    else:
        pt = np.concatenate([np.asarray(x, np.uint8) for x in type_lists])
    return Snapshot(number=number, redshift=redshift, halo_id=np.asarray(halo_ids, np.uint64),
                    npart=np.asarray(npart if npart is not None else counts, np.int32),
                    offsets=offsets, pid=pid, ptype=pt)


# ----------------------------------------------------------------------------------------
# hand-crafted adversarial cases (SURVEY.md section 8c list)

def _ids(number, n):
    return [number * 10**12 + i + 1 for i in range(n)]


def case_threshold_edges(min_part_cmp=10, min_part_halo=30):
    """sum nCommon == minPartCmp (rejected) / +1 (kept); nAllPart == minPartHalo (no token) / +1."""
    base = 1000
    def rng(a, n):
        return list(range(base + a, base + a + n))
    later = [
        rng(0, 100),                     # A0: shares 10 with B0 (rejected) and 11 with B1 (kept)
        rng(1000, min_part_halo),        # A1: no progenitor, nAllPart == minPartHalo  -> vanishes
        rng(2000, min_part_halo + 1),    # A2: no progenitor, nAllPart == minPartHalo+1 -> token (if tracking on)
        rng(3000, 200),                  # A3: plain 1:1 match
    ]
    earlier = [
        rng(0, min_part_cmp) + rng(50000, 40),
        rng(min_part_cmp, min_part_cmp + 1) + rng(51000, 40),
        rng(3000, 180) + rng(52000, 10),
        rng(60000, 25),                  # B3: unrelated small halo
    ]
    return [snapshot_from_lists(20, 0.1, _ids(20, len(earlier)), earlier),
            snapshot_from_lists(21, 0.0, _ids(21, len(later)), later)]


def case_merit_tie():
    """Two candidates with catalogue npart 0 -> both merits are exactly 0.0f -> SortIndexes
    duplicates the first and drops the second (reference src/utils.cpp:438-439); a third,
    ordinary candidate sorts in front.  Also a halo with catalogue npart > 2^24 (float rounding
    of the size ratio)."""
    def rng(a, n):
        return list(range(a, a + n))
    later = [rng(100, 300), rng(5000, 120), rng(9000, 64)]
    earlier = [rng(100, 20) + rng(70000, 5), rng(120, 30) + rng(71000, 5), rng(150, 100) + rng(72000, 5),
               rng(5000, 100), rng(5100, 20) + rng(9000, 60)]
    npart_e = [0, 0, 105, 16777217, 80]
    npart_l = [300, 33554435, 64]
    return [snapshot_from_lists(30, 0.1, _ids(30, 5), earlier, npart=npart_e),
            snapshot_from_lists(31, 0.0, _ids(31, 3), later, npart=npart_l)]


def case_permuted_ids(seed=3):
    """Same memberships as make_chain(seed) but with hashed (unordered) halo IDs: changes iM."""
    return make_chain(seed=seed, n_snaps=5, n_halos=70, id_mode="hashed")


def case_duplicate_membership():
    """One pid listed several times in the same halo and in several halos of both catalogues."""
    later = [[1, 1, 2, 3] + list(range(100, 140)), [2, 3, 3] + list(range(200, 240)), list(range(100, 130))]
    earlier = [[1, 2, 2, 3] + list(range(100, 135)), [1, 3] + list(range(200, 238)) + list(range(120, 135))]
    return [snapshot_from_lists(40, 0.1, _ids(40, 2), earlier), snapshot_from_lists(41, 0.0, _ids(41, 3), later)]


def case_empty_and_ragged():
    """A later catalogue whose halos share nothing with the earlier one, plus 1-particle halos."""
    later = [list(range(10, 60)), [7], list(range(1000, 1100))]
    earlier = [list(range(5000, 5050)), [7], list(range(2000, 2040))]
    return [snapshot_from_lists(50, 0.1, _ids(50, 3), earlier), snapshot_from_lists(51, 0.0, _ids(51, 3), later)]
