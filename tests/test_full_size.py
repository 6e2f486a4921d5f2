"""BASELINE.json configs[1] at full size (1024^3 pair: 1e6 halos, 1e9 bound IDs) through the C ABI.
The oracle cannot hold this pair in reasonable time, so parity is established through properties:
both front ends (partition join, LSD sort) must produce bit-identical raw and clean trees, two runs must agree,
and the trees must satisfy the invariants of the reference's rules (mutual-best links, merit order)."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run(env_extra, args):
    env = dict(os.environ, **env_extra)
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "fullsize_worker.py")] + [str(a) for a in args], capture_output=True, text=True,
                       cwd=ROOT, env=env, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-3000:]
    line = [ln for ln in r.stdout.splitlines() if ln.startswith("FULLSIZE ")][-1]
    return json.loads(line[len("FULLSIZE "):])


@pytest.mark.gpu
@pytest.mark.parametrize("halos,tuples,bits", [(1_000_000, 500_000_000, 30),      # configs[1]: 1024^3 pair on one GPU
                                               (8_000_000, 500_000_000, 30)])     # what one GPU sees of configs[2]: all 8e6 halos of the 2048^3 box
def test_full_size_pair_front_ends_agree(halos, tuples, bits):
    import torch
    if torch.cuda.get_device_properties(0).total_memory < 100e9:
        pytest.skip("needs a 180 GB B200")
    args = (halos, tuples, bits)
    part = run({}, args)
    lsd = run({"METRO_FORCE_LSD": "1"}, args)
    assert part["path"] == 1 and lsd["path"] == 0
    for k in ("n_hits", "n_pairs", "n_trees", "raw", "clean"):
        assert part[k] == lsd[k], k
    for d in (part, lsd):
        # (the sums of the forward and backward raw trees need not agree: the reference's SortIndexes tie quirk duplicates one
        #  candidate and drops another inside a run of equal merits, src/utils.cpp:429-449, frequent among small halos)
        assert d["mutual"] and d["merit_sorted"], d
    assert part["n_hits"] > 5e8 and part["n_trees"] > 0.9 * halos


@pytest.mark.gpu
def test_finer_bucket_retry_gives_the_same_trees():
    """A bucket whose hit staging overflows (catalogues with more memberships per particle than the default bucket size
    provides for) makes the pair retry with buckets half the size instead of dropping to the LSD path; same trees."""
    import torch
    if torch.cuda.get_device_properties(0).total_memory < 100e9:
        pytest.skip("needs a 180 GB B200")
    args = (1_000_000, 500_000_000, 30)
    normal = run({}, args)
    retried = run({"METRO_TEST_WCAP": "150", "METRO_DEBUG": "1"}, args)
    assert retried["path"] == 1
    for k in ("n_hits", "n_pairs", "n_trees", "raw", "clean"):
        assert normal[k] == retried[k], k
