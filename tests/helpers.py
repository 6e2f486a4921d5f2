"""Shared helpers of the GPU parity tests."""
from __future__ import annotations

import numpy as np

import metrocpp_b200 as mb
from oracle import binding


def to_oracle_catalogue(c: mb.HostCatalogue) -> binding.Catalogue:
    typ = c.type if c.type is not None else np.ones(c.n_tuples, np.uint8)
    return binding.Catalogue(np.ascontiguousarray(c.halo_id, np.uint64), np.ascontiguousarray(c.npart, np.int32),
                             np.ascontiguousarray(c.n_orphan_steps, np.int32), np.ascontiguousarray(c.is_token, np.uint8),
                             np.ascontiguousarray(c.pid, np.uint64), np.ascontiguousarray(c.halo, np.uint32),
                             np.ascontiguousarray(typ, np.uint8))


def from_oracle_catalogue(c: binding.Catalogue, noptype: bool) -> mb.HostCatalogue:
    return mb.HostCatalogue(c.halo_id, c.npart, c.n_orphan_steps, c.is_token, c.pid, c.halo, None if noptype else c.type)


def compare_result_with_oracle(res: mb.MatchResult, ores: binding.Result) -> None:
    """Bit-exact: raw trees (order, counts per type, merit bits), clean trees, tokens."""
    for d in range(2):
        np.testing.assert_array_equal(res.raw_off[d], ores.raw_off[d], err_msg=f"raw_off[{d}]")
        np.testing.assert_array_equal(res.raw_prog[d], ores.raw_prog[d], err_msg=f"raw_prog[{d}]")
        np.testing.assert_array_equal(res.raw_ncommon[d], ores.raw_ncommon[d], err_msg=f"raw_ncommon[{d}]")
        np.testing.assert_array_equal(res.raw_merit[d], ores.raw_merit[d], err_msg=f"raw_merit[{d}]")
    np.testing.assert_array_equal(res.tree_main, ores.tree_main, err_msg="tree_main")
    np.testing.assert_array_equal(res.tree_orphan, ores.tree_orphan, err_msg="tree_orphan")
    np.testing.assert_array_equal(res.tree_off, ores.tree_off, err_msg="tree_off")
    np.testing.assert_array_equal(res.clean_prog, ores.clean_prog, err_msg="clean_prog")
    np.testing.assert_array_equal(res.clean_ncommon, ores.clean_ncommon, err_msg="clean_ncommon")
    np.testing.assert_array_equal(res.token_halo, ores.token_halo, err_msg="token_halo")
    assert res.info["n_hits"] == ores.n_hits
    assert res.info["n_pairs"] == ores.n_pairs
    assert res.info["n_untracked"] == ores.n_untracked


def flavour_settings(flavour: str):
    nptypes = 6 if flavour == "zoom_t6" else 2
    noptype = flavour != "zoom_t6"
    fl = mb.FLAVOUR_GATHER if flavour == "gather" else mb.FLAVOUR_ZOOM_DUP
    return nptypes, noptype, fl


def engine_run_chain(eng: mb.MetroEngine, snapshots: list, params: dict, flavour: str, check_oracle: bool = True, on_step=None) -> dict:
    """The reference's main loop (src/main.cpp:154-308) through the C ABI:
    upload -> match -> fetch -> WriteTree text -> shift, one step per older snapshot."""
    nptypes, noptype, fl = flavour_settings(flavour)
    prm = mb.make_params(nptypes, params.get("minPartCmp", 10), params.get("minPartHalo", 30),
                         params.get("facOrphanSteps", 15), params.get("maxOrphanSteps") or 0, fl)
    oprm = binding.Params(nptypes, prm.min_part_cmp, prm.min_part_halo, prm.fac_orphan_steps, prm.max_orphan_steps, fl)
    snaps = sorted(snapshots, key=lambda s: -s.number)
    ocat0 = binding.catalogue_from_snapshot(snaps[0], nptypes, noptype)
    eng.upload(0, prm, from_oracle_catalogue(ocat0, noptype))
    number0 = snaps[0].number
    out = {}
    for s in snaps[1:]:
        ocat1 = binding.catalogue_from_snapshot(s, nptypes, noptype)
        eng.upload(1, prm, from_oracle_catalogue(ocat1, noptype))
        res = eng.find_progenitors_and_clean(prm)
        if on_step is not None:
            on_step(eng, res)
        if check_oracle:
            compare_result_with_oracle(res, binding.match_pair(oprm, ocat0, ocat1))
        out[number0] = binding.mtree_text(ocat0, ocat1, res.tree_main, res.tree_orphan, res.tree_off, res.clean_prog,
                                          res.clean_ncommon)
        eng.shift_halos_parts(prm)
        # the host's view of the new catalogue 0 comes back from the device (token halos appended)
        dev0 = eng.fetch_catalogue(0)
        exp0 = binding.shift(oprm, ocat0, ocat1, res.token_halo)
        if check_oracle:
            np.testing.assert_array_equal(dev0.halo_id, exp0.halo_id)
            np.testing.assert_array_equal(dev0.npart, exp0.npart)
            np.testing.assert_array_equal(dev0.n_orphan_steps, exp0.n_orphan_steps)
            np.testing.assert_array_equal(dev0.is_token, exp0.is_token)
            assert sorted_tuples(dev0.pid, dev0.halo, dev0.type) == sorted_tuples(exp0.pid, exp0.halo, exp0.type)
        ocat0 = to_oracle_catalogue(dev0)
        number0 = s.number
    return out


def sorted_tuples(pid, halo, typ):
    typ = typ if typ is not None else np.ones(pid.shape[0], np.uint8)
    order = np.lexsort((typ, halo, pid))
    return (pid[order].tobytes(), halo[order].tobytes(), typ[order].tobytes())
