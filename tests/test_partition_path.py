"""The partition-join front end on skewed / adversarial pairs and on the golden chains (GPU)."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
@pytest.mark.parametrize("eager", ["1", "0"])
def test_partition_front_end_cases_vs_oracle(eager):
    """eager=1 (default): passes A and B of catalogue 0 run at the end of its upload, under the upload of catalogue 1
    (stage_partition_eager); eager=0: every pair partitions both catalogues inside metro_match_pair."""
    env = dict(os.environ, METRO_PART_MIN_N="1000", METRO_DEBUG="1", METRO_NO_EAGER="0" if eager == "1" else "1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "partition_cases_worker.py")], capture_output=True, text=True,
                       cwd=ROOT, env=env, timeout=900)
    assert r.returncode == 0 and "PARTITION_CASES_OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
    # the two adversarial inputs took the fallback and said why
    assert r.stderr.count("partition join attempt overflowed") >= 2, r.stderr[-2000:]
