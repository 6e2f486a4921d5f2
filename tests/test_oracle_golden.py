"""The C restatement (oracle/) against the reference's own output.

* golden fixtures: .mtree text produced by the UNMODIFIED reference binaries
  (tests/golden/make_golden.py), byte-for-byte;
* merit known answers measured on the compiled reference;
* where /root/reference and oracle/_ref exist (build container only), fresh seeded chains are
  pushed through the live reference as well.
"""
import numpy as np
import pytest

from oracle import binding, run_ref
from tests import cases, golden_io

MERIT_KATS = [  # (nComm, nPH0, nPH1, iM) -> fp32 bits
    ((725, 744, 788, 0), 0x46227446), ((362, 770, 813, 1), 0x45AA5D2F), ((750, 1556, 813, 0), 0x4448F4D7),
    ((1449, 1498, 1556, 0), 0x46E687CF), ((11, 20, 20, 3), 0x4489810E),
    ((100000, 16777217, 33554435, 2), 0x47BF7C94), ((5, 1000, 0, 0), 0x00000000),
    ((12345, 716384, 650000, 7), 0x47D51B0E),
]


@pytest.mark.parametrize("args,bits", MERIT_KATS)
def test_merit_known_answers(args, bits):
    assert binding.lib().oracle_merit_bits(*args) == bits


def test_sort_indexes_tie_quirk():
    # reference probe: merits {5,3,5,7} -> final original indices 3,0,0,1
    import ctypes as C
    m = np.array([5.0, 3.0, 5.0, 7.0], np.float32).view(np.uint32)
    order = np.zeros(4, np.int32)
    binding.lib().oracle_sort_by_merit_order(C.c_void_p(m.ctypes.data), 4, C.c_void_p(order.ctypes.data))
    assert order.tolist() == [3, 0, 0, 1]


@pytest.mark.parametrize("name", golden_io.names())
def test_oracle_matches_golden(name):
    snaps, params, mtree = golden_io.load(name)
    for flavour, expect in mtree.items():
        got = binding.run_chain(snaps, params, flavour=flavour)
        assert sorted(got) == sorted(expect)
        for number in expect:
            assert got[number] == expect[number], f"{name}/{flavour}/step {number}"


@pytest.mark.skipif(not run_ref.available("zoom"), reason="reference not built here (GPU box)")
@pytest.mark.parametrize("seed", [101, 102, 103])
@pytest.mark.parametrize("flavour,track", [("zoom", None), ("zoom", 4), ("gather", 4), ("zoom_t6", 4)])
def test_oracle_matches_live_reference(seed, flavour, track):
    ch = cases.make_chain(seed=seed, n_snaps=5, n_halos=70, multi_species=(flavour == "zoom_t6"),
                          id_mode="hashed" if seed == 103 else "snap")
    prm = dict(minPartHalo=30, minPartCmp=10, facOrphanSteps=15, maxOrphanSteps=track)
    ref = run_ref.run_reference(ch, prm, flavour=flavour)
    got = binding.run_chain(ch, prm, flavour=flavour)
    assert ref["returncode"] == 0 and sorted(ref["mtree"]) == sorted(got)
    for number in got:
        assert got[number] == ref["mtree"][number], (seed, flavour, number)
