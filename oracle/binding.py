"""ctypes binding of the C restatement (oracle/_ref/libmetro_oracle.so).  TEST INFRASTRUCTURE ONLY."""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass

import numpy as np

from . import ahf_io

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "_ref", "libmetro_oracle.so")

FLAVOUR_GATHER, FLAVOUR_ZOOM_DUP = 0, 1


class Params(C.Structure):
    _fields_ = [("nptypes", C.c_int32), ("min_part_cmp", C.c_int32), ("min_part_halo", C.c_int32),
                ("fac_orphan_steps", C.c_int32), ("max_orphan_steps", C.c_int32), ("flavour", C.c_int32)]


class CSnapshot(C.Structure):
    _fields_ = [("n_halos", C.c_int64), ("halo_id", C.c_void_p), ("npart", C.c_void_p),
                ("n_orphan_steps", C.c_void_p), ("is_token", C.c_void_p), ("n_tuples", C.c_int64),
                ("pid", C.c_void_p), ("halo", C.c_void_p), ("type", C.c_void_p)]


class CResult(C.Structure):
    _fields_ = [("raw_count", C.c_int64 * 2), ("raw_off", C.c_void_p * 2), ("raw_prog", C.c_void_p * 2),
                ("raw_ncommon", C.c_void_p * 2), ("raw_merit", C.c_void_p * 2),
                ("n_trees", C.c_int64), ("tree_main", C.c_void_p), ("tree_orphan", C.c_void_p),
                ("tree_off", C.c_void_p), ("clean_count", C.c_int64), ("clean_prog", C.c_void_p),
                ("clean_ncommon", C.c_void_p), ("n_tokens", C.c_int64), ("token_halo", C.c_void_p),
                ("n_untracked", C.c_int64), ("n_hits", C.c_int64), ("n_pairs", C.c_int64)]


_lib = None


def build() -> None:
    subprocess.run(["make", "-s", "-C", HERE, "port"], check=True)


def lib():
    global _lib
    if _lib is None:
        if not os.path.isfile(LIB_PATH):
            build()
        _lib = C.CDLL(LIB_PATH)
        _lib.oracle_match_pair.restype = C.c_int
        _lib.oracle_merit_bits.restype = C.c_uint32
        _lib.oracle_merit_bits.argtypes = [C.c_int32] * 4
    return _lib


@dataclass
class Catalogue:
    """Device-shaped catalogue: what the C-ABI's snapshot_upload takes."""
    halo_id: np.ndarray        # uint64[H]
    npart: np.ndarray          # int32[H, T]
    n_orphan_steps: np.ndarray  # int32[H]
    is_token: np.ndarray       # uint8[H]
    pid: np.ndarray            # uint64[N]
    halo: np.ndarray           # uint32[N]
    type: np.ndarray           # uint8[N]

    @property
    def n_halos(self):
        return int(self.halo_id.shape[0])

    @property
    def n_tuples(self):
        return int(self.pid.shape[0])

    def c_struct(self) -> CSnapshot:
        for name, dt in (("halo_id", np.uint64), ("npart", np.int32), ("n_orphan_steps", np.int32),
                         ("is_token", np.uint8), ("pid", np.uint64), ("halo", np.uint32), ("type", np.uint8)):
            a = getattr(self, name)
            assert a.dtype == dt and a.flags.c_contiguous, name
        s = CSnapshot()
        s.n_halos, s.n_tuples = self.n_halos, self.n_tuples
        for name in ("halo_id", "npart", "n_orphan_steps", "is_token", "pid", "halo", "type"):
            setattr(s, name, getattr(self, name).ctypes.data)
        return s


def catalogue_from_snapshot(s: ahf_io.Snapshot, nptypes: int, noptype: bool) -> Catalogue:
    """What IOSettings::ReadHalos/ReadParticles build (reference src/IOSettings.cpp:599-689,
    961-989): nPart[1] = catalogue npart, every other type count 0; with -DNOPTYPE all types = 1."""
    H = s.n_halos
    npart = np.zeros((H, nptypes), np.int32)
    npart[:, 1] = s.npart
    types = np.ones(s.n_tuples, np.uint8) if noptype else np.ascontiguousarray(s.ptype, np.uint8)
    return Catalogue(np.ascontiguousarray(s.halo_id, np.uint64), npart, np.zeros(H, np.int32),
                     np.zeros(H, np.uint8), np.ascontiguousarray(s.pid, np.uint64),
                     s.halo_index_per_tuple(), types)


@dataclass
class Result:
    raw_off: list
    raw_prog: list
    raw_ncommon: list
    raw_merit: list
    tree_main: np.ndarray
    tree_orphan: np.ndarray
    tree_off: np.ndarray
    clean_prog: np.ndarray
    clean_ncommon: np.ndarray
    token_halo: np.ndarray
    n_untracked: int
    n_hits: int
    n_pairs: int


def _arr(ptr, n, dtype, shape=None):
    if n == 0:
        a = np.zeros(0, dtype)
    else:
        a = np.ctypeslib.as_array(C.cast(ptr, C.POINTER(np.ctypeslib.as_ctypes_type(dtype))), shape=(n,)).copy()
    return a.reshape(shape) if shape is not None else a


def match_pair(prm: Params, cat0: Catalogue, cat1: Catalogue) -> Result:
    r = CResult()
    s0, s1 = cat0.c_struct(), cat1.c_struct()
    rc = lib().oracle_match_pair(C.byref(prm), C.byref(s0), C.byref(s1), C.byref(r))
    if rc != 0:
        raise RuntimeError(f"oracle_match_pair failed: {rc}")
    T = prm.nptypes
    H = (cat0.n_halos, cat1.n_halos)
    res = Result(
        raw_off=[_arr(r.raw_off[d], H[d] + 1, np.int64) for d in range(2)],
        raw_prog=[_arr(r.raw_prog[d], r.raw_count[d], np.int32) for d in range(2)],
        raw_ncommon=[_arr(r.raw_ncommon[d], r.raw_count[d] * T, np.int32, (-1, T)) for d in range(2)],
        raw_merit=[_arr(r.raw_merit[d], r.raw_count[d], np.uint32) for d in range(2)],
        tree_main=_arr(r.tree_main, r.n_trees, np.int32),
        tree_orphan=_arr(r.tree_orphan, r.n_trees, np.uint8),
        tree_off=_arr(r.tree_off, r.n_trees + 1, np.int64),
        clean_prog=_arr(r.clean_prog, r.clean_count, np.int32),
        clean_ncommon=_arr(r.clean_ncommon, r.clean_count * T, np.int32, (-1, T)),
        token_halo=_arr(r.token_halo, r.n_tokens, np.int32),
        n_untracked=int(r.n_untracked), n_hits=int(r.n_hits), n_pairs=int(r.n_pairs))
    lib().oracle_result_free(C.byref(r))
    return res


def shift(prm: Params, cat0: Catalogue, cat1: Catalogue, token_halo: np.ndarray) -> Catalogue:
    """ShiftHalosPartsGrids restated (the C routine needs only the token list)."""
    r = CResult()
    tok = np.ascontiguousarray(token_halo, np.int32)
    r.n_tokens = tok.shape[0]
    r.token_halo = tok.ctypes.data
    s0, s1 = cat0.c_struct(), cat1.c_struct()
    nh, nt = C.c_int64(), C.c_int64()
    lib().oracle_shift_sizes(C.byref(prm), C.byref(s0), C.byref(s1), C.byref(r), C.byref(nh), C.byref(nt))
    T = prm.nptypes
    out = Catalogue(np.zeros(nh.value, np.uint64), np.zeros((nh.value, T), np.int32),
                    np.zeros(nh.value, np.int32), np.zeros(nh.value, np.uint8),
                    np.zeros(nt.value, np.uint64), np.zeros(nt.value, np.uint32), np.zeros(nt.value, np.uint8))
    lib().oracle_shift_fill(C.byref(prm), C.byref(s0), C.byref(s1), C.byref(r),
                            C.c_void_p(out.halo_id.ctypes.data), C.c_void_p(out.npart.ctypes.data),
                            C.c_void_p(out.n_orphan_steps.ctypes.data), C.c_void_p(out.is_token.ctypes.data),
                            C.c_void_p(out.pid.ctypes.data), C.c_void_p(out.halo.ctypes.data),
                            C.c_void_p(out.type.ctypes.data))
    return out


def mtree_text(cat0: Catalogue, cat1: Catalogue, tree_main, tree_orphan, tree_off, clean_prog,
               clean_ncommon, header: bool = True) -> str:
    """IOSettings::WriteTree (reference src/IOSettings.cpp:1113-1132): type-summed numbers only."""
    out = [ahf_io.MTREE_HEADER] if header else []
    np0 = cat0.npart.sum(axis=1)
    np1 = cat1.npart.sum(axis=1)
    ncs = clean_ncommon.sum(axis=1) if clean_ncommon.shape[0] else np.zeros(0, np.int64)
    for k in range(tree_main.shape[0]):
        a = int(tree_main[k])
        lo, hi = int(tree_off[k]), int(tree_off[k + 1])
        out.append(f"{int(cat0.halo_id[a])} {int(np0[a])} {hi - lo} {int(tree_orphan[k])}\n")
        for c in range(lo, hi):
            b = int(clean_prog[c])
            if b < 0:
                out.append(f"{int(np0[a])} {int(cat0.halo_id[a])} {int(ncs[c])}\n")
            else:
                out.append(f"{int(np1[b])} {int(cat1.halo_id[b])} {int(ncs[c])}\n")
    return "".join(out)


def run_chain(snapshots: list, params: dict, flavour: str = "zoom") -> dict:
    """Oracle equivalent of oracle.run_ref.run_reference: {later_snapshot_number: mtree text}."""
    nptypes = 6 if flavour == "zoom_t6" else 2
    noptype = flavour != "zoom_t6"
    prm = Params(nptypes, params.get("minPartCmp", 10), params.get("minPartHalo", 30),
                 params.get("facOrphanSteps", 15), params.get("maxOrphanSteps") or 0,
                 FLAVOUR_GATHER if flavour == "gather" else FLAVOUR_ZOOM_DUP)
    snaps = sorted(snapshots, key=lambda s: -s.number)
    cat0 = catalogue_from_snapshot(snaps[0], nptypes, noptype)
    number0 = snaps[0].number
    out = {}
    for s in snaps[1:]:
        cat1 = catalogue_from_snapshot(s, nptypes, noptype)
        res = match_pair(prm, cat0, cat1)
        out[number0] = mtree_text(cat0, cat1, res.tree_main, res.tree_orphan, res.tree_off,
                                  res.clean_prog, res.clean_ncommon)
        cat0 = shift(prm, cat0, cat1, res.token_halo)
        number0 = s.number
    return out
