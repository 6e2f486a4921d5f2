"""Worker of tests/test_full_size.py: the 1024^3 synthetic pair (BASELINE.json configs[1]) through the C ABI.
Prints one line of digests and invariants; run once per front end (METRO_FORCE_LSD=1 for the LSD path)."""
import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import metrocpp_b200 as mb  # noqa: E402


def digest(*arrays):
    h = hashlib.sha1()
    for a in arrays:
        h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest()


def main():
    halos, tuples, bits = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
    prm = mb.make_params(nptypes=2, min_part_cmp=10, min_part_halo=100, fac_orphan_steps=100, max_orphan_steps=15, flavour=mb.FLAVOUR_GATHER)
    spec = mb.SynthSpec(seed=20261019, n_halos=halos, n_tuples=tuples, pid_bits=bits, shard=0, n_shards=1)
    eng = mb.MetroEngine(0)
    eng.synth_pair_device(prm, spec)
    out = {}
    for rep in range(2):
        info = eng.match_pair(prm)
        res = eng.fetch_result(info, raw=True)
        d = {"path": eng.timings()["path"], "n_hits": info["n_hits"], "n_pairs": info["n_pairs"], "n_trees": info["n_trees"],
             "raw": digest(res.raw_off[0], res.raw_prog[0], res.raw_ncommon[0], res.raw_merit[0], res.raw_off[1], res.raw_prog[1],
                           res.raw_ncommon[1], res.raw_merit[1]),
             "clean": digest(res.tree_main, res.tree_orphan, res.tree_off, res.clean_prog, res.clean_ncommon, res.token_halo)}
        # invariants that do not need an oracle
        f_tot, b_tot = int(res.raw_ncommon[0].sum(dtype=np.int64)), int(res.raw_ncommon[1].sum(dtype=np.int64))
        d["fwd_eq_bwd"] = f_tot == b_tot                         # both raw trees hold the same pairs above the threshold
        d["le_hits"] = f_tot <= info["n_hits"]
        # every clean link is mutual: the main halo is the best descendant (first backward entry) of each kept progenitor
        best = np.full(info["n_halos"][1], -1, np.int64)
        has = np.diff(res.raw_off[1]) > 0
        best[has] = res.raw_prog[1][res.raw_off[1][:-1][has]]
        main_of = np.repeat(res.tree_main, np.diff(res.tree_off))
        real = res.clean_prog >= 0
        d["mutual"] = bool(np.all(best[res.clean_prog[real]] == main_of[real]))
        # merit order inside every forward tree is non-increasing
        m = res.raw_merit[0].view(np.float32)
        seg = np.repeat(np.arange(info["n_halos"][0]), np.diff(res.raw_off[0]))
        same = seg[1:] == seg[:-1]
        d["merit_sorted"] = bool(np.all(m[1:][same] <= m[:-1][same]))
        out[rep] = d
    assert out[0] == out[1], "two runs on the same input differ"          # determinism (record order is not deterministic, results must be)
    print("FULLSIZE " + json.dumps(out[0]), flush=True)
    eng.close()


if __name__ == "__main__":
    main()
