"""BASELINE.json configs[3] and [4] at the oracle-sized variants SURVEY.md section 8d names, every step of the chain compared in
full with the C oracle (raw trees, merit bits, clean trees, tokens, shifted catalogue) through the C ABI:

* chain20  full-box flavour (GATHER_TREES token rule), 20 distinct snapshots, 5e3 halos / 2e6 tuples, tokens on
           (config/fullbox_17_11.cfg: minPartHalo=100 minPartCmp=10 facOrphanSteps=100 maxOrphanSteps=15)
* lgf54    multi-species ZOOM flavour (NPTYPES=6, types {0, 1, 4}, 2 % of the gas turns into stars per step), 54 distinct snapshots,
           2e3 halos / 1e6 tuples (config/lgf_gas.cfg: minPartHalo=30 minPartCmp=10 facOrphanSteps=50, maxOrphanSteps unset)

The snapshots come from the numpy chain generator that `bench.py --workload chain20|lgf54` times at full size."""
import os

import pytest

import metrocpp_b200 as mb
from oracle import synth_np
from tests.helpers import engine_run_chain


@pytest.mark.gpu
@pytest.mark.parametrize("name,n_snaps,halos,tuples,multi,flavour,params", [
    ("chain20", 20, 5000, 2_000_000, False, "gather", dict(minPartHalo=100, minPartCmp=10, facOrphanSteps=100, maxOrphanSteps=15)),
    ("lgf54", 54, 2000, 1_000_000, True, "zoom_t6", dict(minPartHalo=30, minPartCmp=10, facOrphanSteps=50, maxOrphanSteps=None)),
])
def test_chain_workload_vs_oracle(name, n_snaps, halos, tuples, multi, flavour, params, monkeypatch):
    monkeypatch.setenv("METRO_PART_MIN_N", "100000")          # the partition-join front end (with resident buckets) at this size
    snaps = synth_np.synth_chain(20261020, n_snaps, halos, tuples, multi_species=multi)
    eng = mb.MetroEngine(0)
    seen = {"resident": 0, "partition": 0, "tokens": 0}

    def on_step(e, res):
        t = e.timings()
        seen["resident"] += t["resident"] & 1; seen["partition"] += t["path"] == 1; seen["tokens"] += res.info["n_tokens"]
    out = engine_run_chain(eng, snaps, params, flavour, check_oracle=True, on_step=on_step)
    eng.close()
    assert len(out) == n_snaps - 1
    assert seen["partition"] == n_snaps - 1, seen
    assert seen["resident"] >= (n_snaps - 1) // 2, seen       # most steps of a chain re-use the previous pair's buckets
    if flavour == "gather":
        assert seen["tokens"] > 0, seen                       # orphan tracking was exercised
