"""Generate the golden fixtures under tests/golden/ by running the UNMODIFIED reference
(oracle/_ref/MetroCPP_*, built from /root/reference by `make -C oracle ref`).

Only runs where /root/reference exists (the build container).  Each fixture is one
compressed .npz holding the input snapshots AND the reference's .mtree text per step, so the
tests need neither the reference nor numpy's RNG stream to stay stable.

    python -m tests.golden.make_golden
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import run_ref  # noqa: E402
from tests import cases  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))

ZOOM_LAPTOP = dict(minPartHalo=30, minPartCmp=10, facOrphanSteps=15, maxOrphanSteps=None)   # config/zoom_laptop.cfg
LGF_GAS = dict(minPartHalo=30, minPartCmp=10, facOrphanSteps=50, maxOrphanSteps=None)        # config/lgf_gas.cfg
FULLBOX = dict(minPartHalo=100, minPartCmp=10, facOrphanSteps=100, maxOrphanSteps=15)        # config/fullbox_17_11.cfg
TRACK5 = dict(minPartHalo=30, minPartCmp=10, facOrphanSteps=15, maxOrphanSteps=5)

FIXTURES = {
    # name: (chain builder, params, flavours)
    "chain4_zoom_untracked": (lambda: cases.make_chain(seed=11, n_snaps=4, n_halos=60), ZOOM_LAPTOP, ["zoom", "gather"]),
    "chain6_tokens": (lambda: cases.make_chain(seed=12, n_snaps=6, n_halos=60), TRACK5, ["zoom", "gather"]),
    "threshold_edges": (lambda: cases.case_threshold_edges(), TRACK5, ["zoom", "gather"]),
    "merit_tie": (lambda: cases.case_merit_tie(), TRACK5, ["zoom", "gather"]),
    "hashed_ids": (lambda: cases.case_permuted_ids(), TRACK5, ["zoom", "gather"]),
    "duplicate_membership": (lambda: cases.case_duplicate_membership(), dict(TRACK5, minPartCmp=2), ["zoom", "gather"]),
    "empty_and_ragged": (lambda: cases.case_empty_and_ragged(), TRACK5, ["zoom", "gather"]),
    "multi_species": (lambda: cases.make_chain(seed=13, n_snaps=5, n_halos=50, multi_species=True), LGF_GAS, ["zoom_t6"]),
    "multi_species_tokens": (lambda: cases.make_chain(seed=14, n_snaps=5, n_halos=50, multi_species=True), TRACK5, ["zoom_t6"]),
    "fullbox_chain20": (lambda: cases.make_chain(seed=15, n_snaps=20, n_halos=40, mean_size=400, p_vanish=0.08,
                                                 min_size=60), FULLBOX, ["gather", "zoom"]),
}


def pack(snaps):
    d = {"numbers": np.array([s.number for s in snaps]), "redshifts": np.array([s.redshift for s in snaps])}
    for s in snaps:
        k = f"s{s.number}_"
        d[k + "halo_id"], d[k + "npart"], d[k + "offsets"] = s.halo_id, s.npart, s.offsets
        d[k + "pid"], d[k + "ptype"] = s.pid, s.ptype
    return d


def main():
    for name, (builder, params, flavours) in FIXTURES.items():
        snaps = builder()
        d = pack(snaps)
        meta = {"params": params, "flavours": flavours, "mtree": {}}
        for fl in flavours:
            r = run_ref.run_reference(snaps, params, flavour=fl)
            assert r["returncode"] == 0 and len(r["mtree"]) == len(snaps) - 1, (name, fl, r["stdout"][-2000:])
            meta["mtree"][fl] = {str(k): v for k, v in r["mtree"].items()}
        d["meta"] = np.frombuffer(json.dumps(meta).encode(), np.uint8)
        path = os.path.join(HERE, name + ".npz")
        np.savez_compressed(path, **d)
        print(f"{name}: {os.path.getsize(path)} bytes, tuples {[s.n_tuples for s in snaps]}")


if __name__ == "__main__":
    main()
