"""Golden digests at the sizes SURVEY.md section 8c items 6 and 7 name, made by running the UNMODIFIED reference
(oracle/_ref/MetroCPP_*) here in the build container.  The inputs are far too big to commit, so each entry of
tests/golden/big_sha.json holds the generator spec, a SHA-256 of the generated input arrays (a drifted generator is then
reported as such, not as a parity failure) and the SHA-256 of every .mtree file the reference wrote.

    python -m tests.golden.make_golden_big            # ~10 minutes of reference CPU time

* config1_pair      one pair at BASELINE.json configs[0] scale: 2e4 halos, 1e7 bound IDs per snapshot, zoom_laptop.cfg settings,
                    ZOOM build; inputs from the library's deterministic host generator (the GPU workload's generator).
* fullbox_chain20_5k  20 snapshots, 5e3 halos, ~2e6 tuples each, config/fullbox_17_11.cfg settings (minPartHalo=100,
                    facOrphanSteps=100, maxOrphanSteps=15), GATHER_TREES build; inputs from tests/cases.make_chain.
"""
from __future__ import annotations

import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "big_sha.json")

ZOOM_LAPTOP = dict(minPartHalo=30, minPartCmp=10, facOrphanSteps=15, maxOrphanSteps=None)
FULLBOX = dict(minPartHalo=100, minPartCmp=10, facOrphanSteps=100, maxOrphanSteps=15)

SPECS = {
    "config1_pair": {"kind": "synth_pair", "seed": 4242, "n_halos": 20_000, "n_tuples": 10_000_000, "pid_bits": 26,
                     "params": ZOOM_LAPTOP, "flavour": "zoom"},
    "fullbox_chain20_5k": {"kind": "make_chain", "seed": 21, "n_snaps": 20, "n_halos": 5000, "mean_size": 400, "p_vanish": 0.03,
                           "min_size": 60, "params": FULLBOX, "flavour": "gather"},
}


def build_snapshots(spec):
    """Spec -> snapshots (ascending number).  Used by this script and by the tests: nothing here needs the reference."""
    if spec["kind"] == "synth_pair":
        import metrocpp_b200 as mb
        from bench import host_catalogue_to_snapshot
        prm = mb.make_params(nptypes=2)
        sp = mb.SynthSpec(seed=spec["seed"], n_halos=spec["n_halos"], n_tuples=spec["n_tuples"], pid_bits=spec["pid_bits"], shard=0, n_shards=1)
        c0, c1 = mb.synth_pair_host(prm, sp, 0), mb.synth_pair_host(prm, sp, 1)
        return [host_catalogue_to_snapshot(c1, 1, 0.100), host_catalogue_to_snapshot(c0, 2, 0.000)]
    from tests import cases
    return cases.make_chain(seed=spec["seed"], n_snaps=spec["n_snaps"], n_halos=spec["n_halos"], mean_size=spec["mean_size"],
                            p_vanish=spec["p_vanish"], min_size=spec["min_size"])


def input_digest(snaps):
    h = hashlib.sha256()
    for s in snaps:
        for a in (s.halo_id, s.npart, s.offsets, s.pid, s.ptype):
            h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest()


def main():
    from oracle import run_ref
    out = json.load(open(OUT)) if os.path.isfile(OUT) else {}
    for name, spec in SPECS.items():
        if len(sys.argv) > 1 and name not in sys.argv[1:]:
            continue
        snaps = build_snapshots(spec)
        r = run_ref.run_reference(snaps, spec["params"], flavour=spec["flavour"], timeout=7200)
        assert r["returncode"] == 0 and len(r["mtree"]) == len(snaps) - 1, (name, r["stdout"][-2000:])
        out[name] = {"spec": spec, "input_sha256": input_digest(snaps), "tuples": [int(s.n_tuples) for s in snaps],
                     "halos": [int(s.n_halos) for s in snaps], "reference_wall_s": round(r["wall_s"], 1),
                     "mtree_sha256": {str(k): hashlib.sha256(v.encode()).hexdigest() for k, v in r["mtree"].items()},
                     "mtree_bytes": {str(k): len(v) for k, v in r["mtree"].items()}}
        print(name, out[name]["tuples"][:3], "reference wall", out[name]["reference_wall_s"], "s", flush=True)
        with open(OUT, "w") as f:
            json.dump(out, f, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
