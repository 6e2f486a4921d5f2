"""Oracle parity at the sizes the partition-join front end is benchmarked at.

The C oracle (oracle/metro_oracle.c, radix sort + merge) holds 1-2e8 tuples in well under a minute, so the CUDA
path is compared with it FULLY (raw trees incl. merit bits, clean trees, tokens - tests/helpers.compare_result_with_oracle),
not through digests, on pairs that exercise what the small fixtures cannot:

* 1e6 halos / 1.2e8 tuples: 20-bit halo field, 489 hit ranges, sampled hit-region capacities;
* 4e6 halos / 2e8 tuples: hit ranges of 4096 halos, the 1024-thread pass-D blocks;
* the same 1e6-halo pair relabelled the way real AHF files come (halo 0 the most massive, tuples grouped by
  halo, CSR upload): the hit-range capacities are sampled from mass-sorted input and must not trip the fallback.

Every case must have taken the partition join (metro_timings.path == 1) at its first attempt."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run(case, env_extra=None):
    env = dict(os.environ, METRO_DEBUG="1", **(env_extra or {}))
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "parity_at_size_worker.py"), case], capture_output=True, text=True,
                       cwd=ROOT, env=env, timeout=1500)
    assert r.returncode == 0 and "PARITY_AT_SIZE_OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
    assert "partition join attempt overflowed" not in r.stderr, r.stderr[-2000:]
    return r.stdout


@pytest.mark.gpu
@pytest.mark.parametrize("case", ["h1e6", "h4e6", "mass_sorted"])
def test_partition_join_vs_oracle_at_size(case):
    import torch
    if torch.cuda.get_device_properties(0).total_memory < 100e9:
        pytest.skip("needs a 180 GB B200")
    out = run(case)
    assert "path=1" in out, out[-1000:]
