"""Parity at the sizes SURVEY.md section 8c items 6 and 7 name, against digests of what the UNMODIFIED reference wrote
(tests/golden/big_sha.json, made by tests/golden/make_golden_big.py in the build container):

* config1_pair        BASELINE configs[0] scale: 2e4 halos, 1e7 bound IDs per snapshot, zoom_laptop.cfg settings, ZOOM build
* fullbox_chain20_5k  20 snapshots x 5e3 halos x ~1.3e6 tuples, fullbox_17_11.cfg settings (tokens on), GATHER_TREES build

CPU: the C oracle reproduces the reference's .mtree files byte for byte (SHA-256) -> the oracle is pinned at these sizes too.
GPU: the C++ host program (AHF text in, CUDA path, .mtree text out) reproduces them as well.
The inputs are regenerated from the committed spec; their own digest is checked first so that a drifted generator is
reported as such."""
import functools
import hashlib
import json
import os
import subprocess

import pytest

from tests.golden import make_golden_big as big

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "metrocpp_b200", "MetroCPP-b200")
DIGESTS = json.load(open(os.path.join(ROOT, "tests", "golden", "big_sha.json")))


@functools.lru_cache(maxsize=2)
def snapshots(name):
    snaps = big.build_snapshots(DIGESTS[name]["spec"])
    assert big.input_digest(snaps) == DIGESTS[name]["input_sha256"], f"the generator of {name} no longer produces the committed input"
    return snaps


def sha(text):
    return hashlib.sha256(text.encode()).hexdigest()


@pytest.mark.parametrize("name", sorted(DIGESTS))
def test_oracle_reproduces_reference_at_size(name):
    from oracle import binding
    d = DIGESTS[name]
    out = binding.run_chain(snapshots(name), d["spec"]["params"], flavour=d["spec"]["flavour"])
    assert {str(k): sha(v) for k, v in out.items()} == d["mtree_sha256"]


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(DIGESTS))
def test_host_program_reproduces_reference_at_size(tmp_path, name):
    from tests.test_host_program import write_case
    d = DIGESTS[name]
    flavour = d["spec"]["flavour"]
    cfg, out = write_case(str(tmp_path), snapshots(name), d["spec"]["params"], flavour, fullbox_names=(flavour == "gather"))
    r = subprocess.run([BIN, cfg], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    got = {}
    for number in d["mtree_sha256"]:
        with open(os.path.join(out, f"test_{int(number):03d}.0.mtree")) as f:
            got[number] = sha(f.read())
    assert got == d["mtree_sha256"]
