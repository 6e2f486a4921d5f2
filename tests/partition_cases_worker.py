"""Worker of tests/test_partition_path.py (run with METRO_PART_MIN_N small so that the partition-join front
end is taken): adversarial and skewed pairs through the C ABI vs the oracle, plus the golden chains.
Prints one line per case: CASE <name> path=<0|1> ok."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import metrocpp_b200 as mb  # noqa: E402
from oracle import binding  # noqa: E402
from tests import golden_io  # noqa: E402
from tests.helpers import compare_result_with_oracle, engine_run_chain, to_oracle_catalogue  # noqa: E402


def catalogue(halo_lists, base_id, rng, nptypes=2):
    """halo_lists: list of uint64 arrays (particle IDs per halo)."""
    H = len(halo_lists)
    sizes = np.array([len(x) for x in halo_lists], np.int64)
    pid = np.concatenate(halo_lists).astype(np.uint64) if H else np.zeros(0, np.uint64)
    halo = np.repeat(np.arange(H, dtype=np.uint32), sizes)
    npart = np.zeros((H, nptypes), np.int32)
    npart[:, 1] = sizes
    perm = rng.permutation(pid.shape[0])           # tuples in arbitrary order: the API does not require grouping
    return mb.HostCatalogue(np.arange(H, dtype=np.uint64) + np.uint64(base_id), npart, np.zeros(H, np.int32), np.zeros(H, np.uint8),
                            pid[perm], halo[perm], None)


def split(rng, ids, n_halos, alpha=1.0):
    """Partition ids into n_halos lists with power-law-ish sizes."""
    w = rng.random(n_halos) ** (-alpha)
    cuts = np.floor(np.cumsum(w) / w.sum() * len(ids)).astype(np.int64)
    cuts[-1] = len(ids)
    return np.split(ids, cuts[:-1])


def run_case(eng, name, c0, c1, prm, expect_path=None, expect_resident=None):
    eng.upload(0, prm, c0)          # large enough for the partition path: passes A and B of catalogue 0 start here (eager partition)
    eng.upload(1, prm, c1)
    res = eng.find_progenitors_and_clean(prm)
    path, resident = eng.timings()["path"], eng.timings()["resident"]
    oprm = binding.Params(prm.nptypes, prm.min_part_cmp, prm.min_part_halo, prm.fac_orphan_steps, prm.max_orphan_steps, prm.flavour)
    compare_result_with_oracle(res, binding.match_pair(oprm, to_oracle_catalogue(c0), to_oracle_catalogue(c1)))
    if expect_path is not None:
        assert path == expect_path, (name, path, expect_path)
    if expect_resident is not None:                # bit 0: catalogue 0's buckets resident; bit 1: catalogue 1 partitioned under its upload
        assert resident == expect_resident, (name, resident, expect_resident)
    print(f"CASE {name} path={path} eager={resident} hits={res.info['n_hits']} pairs={res.info['n_pairs']} ok", flush=True)


def main():
    rng = np.random.default_rng(2026)
    prm = mb.make_params(nptypes=2, max_orphan_steps=5, flavour=mb.FLAVOUR_GATHER)
    eng = mb.MetroEngine(0)
    n = 400_000
    ids = (rng.permutation(1 << 22)[:n].astype(np.uint64) + np.uint64(1))

    # 1. plain pair on the partition path: 90 % of every halo survives into a permuted earlier catalogue
    later = split(rng, ids, 800)
    earlier = [x[: int(0.9 * len(x))] for x in later]
    order = rng.permutation(len(earlier))
    run_case(eng, "plain", catalogue(later, 10**6, rng), catalogue([earlier[i] for i in order], 2 * 10**6, rng), prm, expect_path=1,
             expect_resident=0 if os.environ.get("METRO_NO_EAGER") == "1" else 3)

    # 1b. hashed 64-bit particle IDs (pid + snap + halo + type do not fit one 64-bit key): the wide-key front end; the
    #     trees must be those of the same pair with narrow IDs (only the equality of IDs matters)
    wide = lambda lists: [x * np.uint64(0x9E3779B97F4A7C15) + np.uint64(1 << 63) for x in lists]     # odd multiplier: a bijection on u64
    with np.errstate(over="ignore"):
        run_case(eng, "wide_ids", catalogue(wide(later), 10**6, rng), catalogue(wide([earlier[i] for i in order]), 2 * 10**6, rng), prm, expect_path=2)

    # 2. one dominant halo (70 % of all tuples): one hit range is far above the mean and gets sliced
    big = [ids[: int(0.7 * n)]] + split(rng, ids[int(0.7 * n):], 300)
    big1 = [x[: int(0.95 * len(x))] for x in big]
    run_case(eng, "dominant_halo", catalogue(big, 10**6, rng), catalogue(big1, 2 * 10**6, rng), prm, expect_path=1)

    # 3. earlier catalogue 4x larger than the later one (fresh particles in extra halos, IDs beyond the later catalogue's range): the
    #    eagerly built buckets of catalogue 0 do not fit the pair and are dropped
    extra = (np.arange(3 * n, dtype=np.uint64) + np.uint64(1 << 23))
    run_case(eng, "unbalanced", catalogue(later, 10**6, rng), catalogue(earlier + split(rng, extra, 1500), 2 * 10**6, rng), prm, expect_path=1,
             expect_resident=0)

    # 4. nesting: every particle of the earlier catalogue is a member of 3 halos (3 hits per catalogue-0 tuple): stays on the
    #    partition path (the hit ranges of a small catalogue carry enough slack), same answer
    nest = earlier + [np.concatenate(earlier[i::40]) for i in range(40)] + [np.concatenate(earlier[i::7]) for i in range(7)]
    run_case(eng, "triple_membership", catalogue(later, 10**6, rng), catalogue(nest, 2 * 10**6, rng), prm, expect_path=1)

    # 4b. deep nesting: every particle is a member of 21 earlier halos (21 hits per catalogue-0 tuple over whole ranges, 20 chain
    #     nodes per table slot => chain nodes / hit regions overflow => finer buckets do not cure it => LSD fallback, same answer)
    deep = list(earlier)
    for k in (3, 5, 7, 11, 13, 17, 19, 23, 29, 31, 37, 41, 43, 47, 53, 59, 61, 67, 71, 73):
        deep += [np.concatenate(earlier[i::k]) for i in range(k)]
    run_case(eng, "deep_nesting", catalogue(later, 10**6, rng), catalogue(deep, 2 * 10**6, rng), prm, expect_path=0)

    # 5. one particle ID listed in 200 earlier halos and 50 later halos (10^4 hits from one ID: the bucket table's
    #    last-resort list overflows => LSD fallback)
    hot = np.uint64(ids[0])
    later_hot = [np.concatenate([x, [hot]]) if i < 50 and hot not in x else x for i, x in enumerate(later)]
    earlier_hot = [np.concatenate([x, [hot]]) if i < 200 and hot not in x else x for i, x in enumerate(earlier)]
    run_case(eng, "hot_particle", catalogue(later_hot, 10**6, rng), catalogue(earlier_hot, 2 * 10**6, rng), prm, expect_path=0)

    # 6. halos with many small partners (every later halo shares particles with 12 earlier halos): exercises the
    #    overflow table of the pair counter
    frag = []
    for x in later:
        frag += np.array_split(x, 12) if len(x) >= 240 else [x]
    run_case(eng, "many_partners", catalogue(later, 10**6, rng), catalogue(frag, 2 * 10**6, rng), prm, expect_path=1)

    # 6b. one later halo whose particles come from 3500 earlier halos of 14 particles each (3500 candidates above the
    #     threshold in one tree: the long-segment merit sort) next to halos with 40 candidates (warp sort spill-over)
    giant = ids[:49_000]
    pieces = np.array_split(giant, 3500)
    mids = split(rng, ids[49_000:200_000], 30)
    mid_pieces = []
    for x in mids:
        mid_pieces += np.array_split(x, 40) if len(x) >= 40 * 12 else [x]
    rest = split(rng, ids[200_000:], 500)
    run_case(eng, "fragmented_giant", catalogue([giant] + mids + rest, 10**6, rng),
             catalogue(pieces + mid_pieces + [x[: int(0.9 * len(x))] for x in rest], 2 * 10**6, rng), prm)

    # 6c. the same with 6000 earlier halos of 12 particles: more candidates in one tree than the shared-memory sort of the long-segment
    #     kernel holds (4096): its counting path
    giant2 = ids[:72_000]
    pieces2 = np.array_split(giant2, 6000)
    rest2 = split(rng, ids[72_000:], 500)
    run_case(eng, "fragmented_giant_6000", catalogue([giant2] + rest2, 10**6, rng),
             catalogue(pieces2 + [x[: int(0.9 * len(x))] for x in rest2], 2 * 10**6, rng), prm)

    # 7. multi-species, ZOOM flavour
    prm6 = mb.make_params(nptypes=6, max_orphan_steps=3, flavour=mb.FLAVOUR_ZOOM_DUP)
    spec = mb.SynthSpec(seed=11, n_halos=3000, n_tuples=600_000, pid_bits=22, shard=0, n_shards=1)
    c0, c1 = mb.synth_pair_host(prm6, spec, 0), mb.synth_pair_host(prm6, spec, 1)
    run_case(eng, "multi_species_zoom", c0, c1, prm6, expect_path=1)

    # 7b. a chain on the partition path: from the second pair on, the key buckets of catalogue 0 are the resident buckets
    #     of the previous pair's catalogue 1 (+ the token tuples appended by the shift); every step is checked against the
    #     oracle, in both token flavours (ZOOM doubles the token tuples)
    from tests.cases import make_chain
    for flavour, seed in (("gather", 21), ("zoom", 22)):
        chain = make_chain(seed=seed, n_snaps=6, n_halos=1500, mean_size=250, p_vanish=0.03)
        steps = []
        engine_run_chain(eng, chain, {"minPartHalo": 30, "minPartCmp": 10, "facOrphanSteps": 15, "maxOrphanSteps": 3}, flavour,
                         on_step=lambda e, r: steps.append((e.timings()["path"], e.timings()["resident"], r.info["n_tokens"])))
        assert len(steps) == 5 and all(p == 1 for p, _, _ in steps), steps
        # resident from the second pair on (and, with the eager partition, in the first), unless the catalogue outgrows the plan (the
        # ZOOM flavour doubles the token tuples at every orphan step: src/MergerTree.cpp:597-603)
        assert sum(r & 1 for _, r, _ in steps[1:]) >= (3 if flavour == "gather" else 1), steps
        # whenever catalogue 0's buckets were resident, catalogue 1 had been through pass A under its upload (unless switched off)
        assert all(r in ((0, 1) if os.environ.get("METRO_NO_EAGER") == "1" else (0, 3)) for _, r, _ in steps), steps
        assert sum(t for _, _, t in steps) > 0, steps                                  # token tuples were appended to resident buckets
        print(f"CASE resident_chain_{flavour} steps={steps} ok", flush=True)

    # 8. the golden chains (tokens, shifts, ties, thresholds) with the partition path wherever it applies
    for name in golden_io.names():
        snaps, params, mtree = golden_io.load(name)
        for flavour, expect in mtree.items():
            got = engine_run_chain(eng, snaps, params, flavour)
            assert got == expect, f"{name}/{flavour}"
        print(f"CASE golden_{name} ok", flush=True)
    eng.close()
    print("PARTITION_CASES_OK")


if __name__ == "__main__":
    main()
