"""GPU parity tests: the CUDA path through the C ABI against the oracle and the golden fixtures."""
import numpy as np
import pytest

import metrocpp_b200 as mb
from oracle import binding
from tests import golden_io
from tests.helpers import (compare_result_with_oracle, engine_run_chain, from_oracle_catalogue, sorted_tuples,
                           to_oracle_catalogue)
from tests.test_oracle_golden import MERIT_KATS

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    e = mb.MetroEngine(0)
    yield e
    e.close()


def test_device_merit_known_answers(eng):
    args = np.array([a for a, _ in MERIT_KATS], np.int32)
    bits = eng.merit_bits(args[:, 0], args[:, 1], args[:, 2], args[:, 3])
    assert bits.tolist() == [b for _, b in MERIT_KATS]


def test_device_merit_random_vs_oracle(eng):
    rng = np.random.default_rng(5)
    n = 200_000
    ph0 = rng.integers(1, 40_000_000, n).astype(np.int32)
    ph1 = np.where(rng.random(n) < 0.5, ph0 + rng.integers(-50, 50, n), rng.integers(0, 40_000_000, n)).astype(np.int32)
    ph1 = np.maximum(ph1, 0)
    ncm = rng.integers(1, 2_000_000, n).astype(np.int32)
    im = rng.integers(0, 3000, n).astype(np.int32)
    got = eng.merit_bits(ncm, ph0, ph1, im)
    lib = binding.lib()
    exp = np.array([lib.oracle_merit_bits(int(a), int(b), int(c), int(d)) for a, b, c, d in zip(ncm[:20000], ph0, ph1, im)], np.uint32)
    np.testing.assert_array_equal(got[:20000], exp)


@pytest.mark.parametrize("n,begin,end,vals", [(0, 0, 64, False), (1, 0, 64, True), (1000, 0, 64, True), (6144, 3, 35, False),
                                                (6145, 0, 8, True), (100_003, 17, 47, True), (3_000_017, 0, 64, False),
                                                (2_000_000, 24, 54, True), (1_500_000, 5, 11, True)])
def test_radix_sort_stable(eng, n, begin, end, vals):
    import torch
    rng = np.random.default_rng(n + begin)
    keys = rng.integers(0, 1 << 63, size=n, dtype=np.uint64) * np.uint64(2) + rng.integers(0, 2, size=n, dtype=np.uint64)
    if n > 10:
        keys[: n // 3] = keys[n // 2: n // 2 + n // 3]          # many duplicates
    v = np.arange(n, dtype=np.uint32)
    dk = torch.from_numpy(keys.view(np.int64)).cuda()
    dk2 = torch.empty_like(dk)
    dv = torch.from_numpy(v.view(np.int32)).cuda()
    dv2 = torch.empty_like(dv)
    torch.cuda.synchronize()
    eng.sort_u64_device(dk.data_ptr(), dk2.data_ptr(), dv.data_ptr() if vals else None, dv2.data_ptr() if vals else None, n, begin, end)
    out = dk2.cpu().numpy().view(np.uint64)
    digit = (keys >> np.uint64(begin)) & np.uint64((1 << (end - begin)) - 1 if end - begin < 64 else 0xFFFFFFFFFFFFFFFF)
    order = np.argsort(digit, kind="stable")
    np.testing.assert_array_equal(out, keys[order])
    if vals:
        np.testing.assert_array_equal(dv2.cpu().numpy().view(np.uint32), v[order])


@pytest.mark.parametrize("name", golden_io.names())
def test_golden_chain_through_c_abi(eng, name):
    snaps, params, mtree = golden_io.load(name)
    for flavour, expect in mtree.items():
        got = engine_run_chain(eng, snaps, params, flavour)
        assert sorted(got) == sorted(expect)
        for number in expect:
            assert got[number] == expect[number], f"{name}/{flavour}/step {number}"


@pytest.mark.parametrize("nptypes,halos,tuples,bits", [(2, 300, 30_000, 18), (2, 5_000, 1_000_000, 22), (6, 2_000, 400_000, 22)])
def test_synthetic_pair_vs_oracle(eng, nptypes, halos, tuples, bits):
    prm = mb.make_params(nptypes=nptypes, max_orphan_steps=5, flavour=mb.FLAVOUR_GATHER)
    spec = mb.SynthSpec(seed=77, n_halos=halos, n_tuples=tuples, pid_bits=bits, shard=0, n_shards=1)
    eng.synth_pair_device(prm, spec)
    d0, d1 = eng.fetch_catalogue(0), eng.fetch_catalogue(1)
    h0, h1 = mb.synth_pair_host(prm, spec, 0), mb.synth_pair_host(prm, spec, 1)
    for d, h in ((d0, h0), (d1, h1)):            # device generator == host generator
        np.testing.assert_array_equal(d.halo_id, h.halo_id)
        np.testing.assert_array_equal(d.npart, h.npart)
        assert sorted_tuples(d.pid, d.halo, d.type) == sorted_tuples(h.pid, h.halo, h.type)
    res = eng.find_progenitors_and_clean(prm)
    oprm = binding.Params(nptypes, prm.min_part_cmp, prm.min_part_halo, prm.fac_orphan_steps, prm.max_orphan_steps, 0)
    ores = binding.match_pair(oprm, to_oracle_catalogue(d0), to_oracle_catalogue(d1))
    compare_result_with_oracle(res, ores)
    assert res.info["n_trees"] > 0.9 * halos and res.info["n_tokens"] > 0


def test_sharded_synthetic_union_equals_unsharded(eng):
    prm = mb.make_params(nptypes=2)
    full = mb.SynthSpec(seed=9, n_halos=400, n_tuples=60_000, pid_bits=18, shard=0, n_shards=1)
    f0 = mb.synth_pair_host(prm, full, 0)
    parts = [mb.synth_pair_host(prm, mb.SynthSpec(seed=9, n_halos=400, n_tuples=60_000, pid_bits=18, shard=s, n_shards=4), 0)
             for s in range(4)]
    pid = np.concatenate([p.pid for p in parts]); halo = np.concatenate([p.halo for p in parts])
    typ = np.concatenate([p.type for p in parts])
    assert sorted_tuples(pid, halo, typ) == sorted_tuples(f0.pid, f0.halo, f0.type)
    for s, p in enumerate(parts):
        assert np.all((p.pid >> np.uint64(16)) == s)


def test_empty_catalogues_and_errors(eng):
    prm = mb.make_params(nptypes=2, max_orphan_steps=3)
    empty = mb.HostCatalogue(np.zeros(0, np.uint64), np.zeros((0, 2), np.int32), np.zeros(0, np.int32), np.zeros(0, np.uint8),
                             np.zeros(0, np.uint64), np.zeros(0, np.uint32), None)
    one = mb.HostCatalogue(np.array([5], np.uint64), np.array([[0, 40]], np.int32), np.zeros(1, np.int32), np.zeros(1, np.uint8),
                           np.arange(40, dtype=np.uint64), np.zeros(40, np.uint32), None)
    eng.upload(0, prm, one); eng.upload(1, prm, empty)
    res = eng.find_progenitors_and_clean(prm)
    assert res.info["n_hits"] == 0 and res.token_halo.tolist() == [0] and res.clean_prog.tolist() == [-1]
    eng.upload(0, prm, empty); eng.upload(1, prm, one)
    res = eng.find_progenitors_and_clean(prm)
    assert res.info["n_trees"] == 0 and res.info["raw_count"] == [0, 0]
    bad = mb.HostCatalogue(one.halo_id, one.npart, one.n_orphan_steps, one.is_token, one.pid, np.full(40, 3, np.uint32), None)
    with pytest.raises(mb.MetroError):
        eng.upload(0, prm, bad)
    with pytest.raises(mb.MetroError):
        eng.match_pair(mb.make_params(nptypes=2, fac_orphan_steps=0))


def test_partition_join_path_vs_oracle():
    """The partition-join front end (large pairs) must give the same trees as the oracle; a second
    engine forced onto the LSD path must agree bit for bit as well."""
    import os
    import subprocess
    import sys
    code = r'''
import os, sys, numpy as np
sys.path.insert(0, os.getcwd())
import metrocpp_b200 as mb
from oracle import binding
from tests.helpers import compare_result_with_oracle, to_oracle_catalogue
prm = mb.make_params(nptypes=int(sys.argv[1]), max_orphan_steps=5, flavour=mb.FLAVOUR_GATHER)
spec = mb.SynthSpec(seed=123, n_halos=6000, n_tuples=3_000_000, pid_bits=24, shard=0, n_shards=1)
eng = mb.MetroEngine(0)
eng.synth_pair_device(prm, spec)
res = eng.find_progenitors_and_clean(prm)
assert eng.timings()["path"] == int(os.environ.get("EXPECT_PATH", "1")), eng.timings()
d0, d1 = eng.fetch_catalogue(0), eng.fetch_catalogue(1)
oprm = binding.Params(prm.nptypes, prm.min_part_cmp, prm.min_part_halo, prm.fac_orphan_steps, prm.max_orphan_steps, 0)
compare_result_with_oracle(res, binding.match_pair(oprm, to_oracle_catalogue(d0), to_oracle_catalogue(d1)))
print("PATH_OK", eng.timings()["path"], res.info["n_hits"])
'''
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for nptypes in ("2", "6"):
        for env_extra, expect in (({"METRO_PART_MIN_N": "100000"}, "1"), ({"METRO_FORCE_LSD": "1"}, "0")):
            env = dict(os.environ, EXPECT_PATH=expect, **env_extra)
            r = subprocess.run([sys.executable, "-c", code, nptypes], capture_output=True, text=True, cwd=root, env=env, timeout=600)
            assert r.returncode == 0 and "PATH_OK " + expect in r.stdout, r.stdout[-2000:] + r.stderr[-3000:]


@pytest.mark.parametrize("nptypes", [2, 6])
def test_csr_upload_rebuilds_the_halo_index(eng, nptypes):
    """metro_catalogue_upload_csr (tuples grouped by halo + offsets, the locParts layout, with empty halos and a
    ragged tail) must leave exactly the catalogue the per-tuple upload leaves, and match identically."""
    prm = mb.make_params(nptypes=nptypes, max_orphan_steps=4)
    spec = mb.SynthSpec(seed=5, n_halos=900, n_tuples=150_003, pid_bits=20, shard=0, n_shards=1)
    cats = [mb.synth_pair_host(prm, spec, s) for s in range(2)]
    grouped = []
    for c in cats:
        order = np.argsort(c.halo, kind="stable")
        off = np.zeros(c.n_halos + 1, np.int64)
        np.cumsum(np.bincount(c.halo, minlength=c.n_halos), out=off[1:])
        grouped.append((mb.HostCatalogue(c.halo_id, c.npart, c.n_orphan_steps, c.is_token, c.pid[order], c.halo[order],
                                         None if c.type is None else c.type[order]), off))
    for slot, (g, off) in enumerate(grouped):
        eng.upload(slot, prm, g, halo_offset=off)
        back = eng.fetch_catalogue(slot)
        np.testing.assert_array_equal(back.halo, g.halo)
        np.testing.assert_array_equal(back.pid, g.pid)
    res_csr = eng.find_progenitors_and_clean(prm)
    for slot, (g, _) in enumerate(grouped):
        eng.upload(slot, prm, g)
    res = eng.find_progenitors_and_clean(prm)
    for k in ("tree_main", "tree_orphan", "tree_off", "clean_prog", "clean_ncommon", "token_halo"):
        np.testing.assert_array_equal(getattr(res_csr, k), getattr(res, k), err_msg=k)
    oprm = binding.Params(nptypes, prm.min_part_cmp, prm.min_part_halo, prm.fac_orphan_steps, prm.max_orphan_steps, 0)
    compare_result_with_oracle(res_csr, binding.match_pair(oprm, to_oracle_catalogue(grouped[0][0]), to_oracle_catalogue(grouped[1][0])))
    bad = grouped[0][1].copy(); bad[3] = bad[4] + 1
    with pytest.raises(mb.MetroError):
        eng.upload(0, prm, grouped[0][0], halo_offset=bad)


@pytest.mark.gpu
def test_csr32_upload_equals_the_u64_upload(eng):
    """metro_catalogue_upload_csr32 (IDs as 32-bit offsets from a base, widened on the device chunk by chunk) must leave exactly the
    catalogue the 64-bit CSR upload leaves - also across several upload chunks and with a large base - and match identically."""
    prm = mb.make_params(nptypes=2, max_orphan_steps=4)
    spec = mb.SynthSpec(seed=9, n_halos=700, n_tuples=120_001, pid_bits=20, shard=0, n_shards=1)
    base = np.uint64(7_000_000_000_000)
    grouped = []
    for s in range(2):
        c = mb.synth_pair_host(prm, spec, s)
        order = np.argsort(c.halo, kind="stable")
        off = np.zeros(c.n_halos + 1, np.int64)
        np.cumsum(np.bincount(c.halo, minlength=c.n_halos), out=off[1:])
        grouped.append((mb.HostCatalogue(c.halo_id, c.npart, c.n_orphan_steps, c.is_token, c.pid[order] + base, c.halo[order], None), off))
    lo = min(int(g.pid.min()) for g, _ in grouped)
    for slot, (g, off) in enumerate(grouped):
        p32 = (g.pid - np.uint64(lo)).astype(np.uint32)
        eng.upload(slot, prm, mb.HostCatalogue(g.halo_id, g.npart, g.n_orphan_steps, g.is_token, None, None, None), halo_offset=off, pid32=p32, pid_base=lo)
        back = eng.fetch_catalogue(slot)
        np.testing.assert_array_equal(back.pid, g.pid)
        np.testing.assert_array_equal(back.halo, g.halo)
    res32 = eng.find_progenitors_and_clean(prm)
    oprm = binding.Params(2, prm.min_part_cmp, prm.min_part_halo, prm.fac_orphan_steps, prm.max_orphan_steps, prm.flavour)
    compare_result_with_oracle(res32, binding.match_pair(oprm, to_oracle_catalogue(grouped[0][0]), to_oracle_catalogue(grouped[1][0])))
    with pytest.raises(mb.MetroError):
        eng.upload(0, prm, mb.HostCatalogue(grouped[0][0].halo_id, grouped[0][0].npart, None, None, None, None, None), halo_offset=grouped[0][1],
                   pid32=np.zeros(grouped[0][0].n_tuples, np.uint32), pid_base=(1 << 64) - 5)
