"""Worker of tests/test_parity_at_size.py: one big synthetic pair through the C ABI, compared in full with the C oracle.
usage: parity_at_size_worker.py {h1e6|h4e6|mass_sorted}"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import metrocpp_b200 as mb  # noqa: E402
from oracle import binding  # noqa: E402
from tests.helpers import compare_result_with_oracle, to_oracle_catalogue  # noqa: E402

CASES = {              # halos, tuples per catalogue (target), pid bits
    "h1e6": (1_000_000, 50_000_000, 28),
    "h4e6": (4_000_000, 90_000_000, 30),
    "mass_sorted": (1_000_000, 50_000_000, 28),
}


def mass_sorted(c: mb.HostCatalogue):
    """Relabel the halos by descending size and group the tuples by halo: the order of a real AHF catalogue."""
    sizes = np.bincount(c.halo, minlength=c.n_halos)
    order = np.argsort(-sizes, kind="stable")                  # new index -> old index
    new_of_old = np.empty(c.n_halos, np.uint32)
    new_of_old[order] = np.arange(c.n_halos, dtype=np.uint32)
    halo = new_of_old[c.halo]
    grp = np.argsort(halo, kind="stable")
    off = np.zeros(c.n_halos + 1, np.int64)
    np.cumsum(sizes[order], out=off[1:])
    typ = None if c.type is None else c.type[grp]
    return mb.HostCatalogue(c.halo_id[order], c.npart[order], c.n_orphan_steps[order], c.is_token[order], c.pid[grp], halo[grp], typ), off


def main():
    case = sys.argv[1]
    halos, tuples, bits = CASES[case]
    prm = mb.make_params(nptypes=2, min_part_cmp=10, min_part_halo=100, fac_orphan_steps=100, max_orphan_steps=15, flavour=mb.FLAVOUR_GATHER)
    oprm = binding.Params(2, prm.min_part_cmp, prm.min_part_halo, prm.fac_orphan_steps, prm.max_orphan_steps, 0)
    spec = mb.SynthSpec(seed=20261019, n_halos=halos, n_tuples=tuples, pid_bits=bits, shard=0, n_shards=1)
    eng = mb.MetroEngine(0)
    eng.synth_pair_device(prm, spec)
    c0, c1 = eng.fetch_catalogue(0), eng.fetch_catalogue(1)
    c0.type = c1.type = None                                   # NOPTYPE build: every type is 1
    if case == "mass_sorted":
        (c0, off0), (c1, off1) = mass_sorted(c0), mass_sorted(c1)
        eng.upload(0, prm, c0, halo_offset=off0)
        eng.upload(1, prm, c1, halo_offset=off1)
    t = time.time()
    res = eng.find_progenitors_and_clean(prm)
    tm = eng.timings()
    assert tm["path"] == 1, tm
    t_gpu = time.time() - t
    t = time.time()
    o0, o1 = to_oracle_catalogue(c0), to_oracle_catalogue(c1)
    ores = binding.match_pair(oprm, o0, o1)
    t_or = time.time() - t
    compare_result_with_oracle(res, ores)
    print(f"CASE {case} path={tm['path']} halos={c0.n_halos}/{c1.n_halos} tuples={c0.n_tuples}+{c1.n_tuples} hits={res.info['n_hits']} "
          f"pairs={res.info['n_pairs']} trees={res.info['n_trees']} tokens={res.info['n_tokens']} halo_bits={res.info['halo_bits']} "
          f"device={tm['total']:.2f}ms wall={t_gpu:.2f}s oracle={t_or:.1f}s", flush=True)
    if case == "h1e6":
        # the shift at size: token halos appended, tuple multiset identical to the oracle's shift
        eng.shift_halos_parts(prm)
        dev0 = eng.fetch_catalogue(0)
        exp0 = binding.shift(oprm, o0, o1, res.token_halo)
        np.testing.assert_array_equal(dev0.halo_id, exp0.halo_id)
        np.testing.assert_array_equal(dev0.npart, exp0.npart)
        np.testing.assert_array_equal(dev0.n_orphan_steps, exp0.n_orphan_steps)
        np.testing.assert_array_equal(dev0.is_token, exp0.is_token)
        H = np.uint64(dev0.n_halos)
        a = np.sort(dev0.pid * H + dev0.halo.astype(np.uint64))
        b = np.sort(exp0.pid * H + exp0.halo.astype(np.uint64))
        np.testing.assert_array_equal(a, b)
        # and the next pair of the chain on the resident buckets: the old catalogue 1 (now 0, with tokens) against the old catalogue 0's
        # later half as a stand-in earlier snapshot is not meaningful physics, but it is a valid input: compare again in full
        eng.upload(1, prm, c1)
        res2 = eng.find_progenitors_and_clean(prm)
        tm2 = eng.timings()
        ores2 = binding.match_pair(oprm, exp0, o1)
        compare_result_with_oracle(res2, ores2)
        print(f"CASE {case}+shift path={tm2['path']} resident={tm2['resident']} tokens_in={int(dev0.is_token.sum())} hits={res2.info['n_hits']} ok", flush=True)
    print("PARITY_AT_SIZE_OK", flush=True)
    eng.close()


if __name__ == "__main__":
    main()
