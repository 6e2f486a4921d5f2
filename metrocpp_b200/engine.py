"""Python mirror of the reference's matching interface on top of the C ABI.

The reference exposes the hot path as free functions over globals
(reference src/MergerTree.h:85-98, src/utils.h:47):

    FindProgenitors(0,1); FindProgenitors(1,0); CleanTrees(iStep); ShiftHalosPartsGrids()

`MetroEngine` keeps those steps and their order (`find_progenitors_and_clean` = the three
matching calls fused into one device pipeline, `shift_halos_parts` = the shift) with numpy
arrays standing in for locHalos / locMapParts.  It only marshals pointers: all computation is in
libmetro_b200.so; a missing library or GPU raises, nothing falls back to the CPU.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np

from . import _lib
from ._lib import Params, SynthSpec


class MetroError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"metro_b200 error {code}: {msg}")
        self.code = code


@dataclass
class HostCatalogue:
    halo_id: np.ndarray          # uint64[H]
    npart: np.ndarray            # int32[H, T]
    n_orphan_steps: np.ndarray   # int32[H]
    is_token: np.ndarray         # uint8[H]
    pid: np.ndarray              # uint64[N]
    halo: np.ndarray             # uint32[N]
    type: np.ndarray | None      # uint8[N] or None (= all 1, the NOPTYPE build)

    @property
    def n_halos(self):
        return int(self.halo_id.shape[0])

    @property
    def n_tuples(self):
        return int(self.pid.shape[0])


@dataclass
class MatchResult:
    info: dict
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


def make_params(nptypes=2, min_part_cmp=10, min_part_halo=30, fac_orphan_steps=15, max_orphan_steps=0,
                flavour=_lib.FLAVOUR_ZOOM_DUP) -> Params:
    return Params(nptypes, min_part_cmp, min_part_halo, fac_orphan_steps, max_orphan_steps or 0, flavour)


def _ptr(a):
    return None if a is None else a.ctypes.data


class MetroEngine:
    def __init__(self, device: int = 0):
        self.lib = _lib.load()
        h = C.c_void_p()
        rc = self.lib.metro_ctx_create(device, C.byref(h))
        if rc != 0:
            raise MetroError(rc, self.lib.metro_last_error(None).decode())
        self.h = h
        self.nptypes = None

    def close(self):
        if getattr(self, "h", None):
            self.lib.metro_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_stream(self, cuda_stream: int | None):
        """Run on the caller's CUDA stream (e.g. torch.cuda.current_stream().cuda_stream)."""
        self._check(self.lib.metro_ctx_set_stream(self.h, cuda_stream))

    def _check(self, rc):
        if rc != 0:
            raise MetroError(rc, self.lib.metro_last_error(self.h).decode())

    # -- catalogues (locHalos[slot], locMapParts[slot]) ---------------------------------------
    def upload(self, slot: int, prm: Params, cat: HostCatalogue, device_ptrs: tuple | None = None, halo_offset=None, pid32=None, pid_base: int = 0):
        """halo_offset (int64[H+1]): tuples are grouped by halo (the locParts layout); cat.halo is then not sent.
        pid32 (uint32[N], with halo_offset): particle IDs as 32-bit offsets from pid_base; cat.pid is then not sent."""
        T = prm.nptypes
        c = _lib.Catalogue()
        keep = []
        def arr(a, dt, shape=None):
            a = np.ascontiguousarray(a, dt)
            if shape is not None:
                assert a.shape == shape, (a.shape, shape)
            keep.append(a)
            return a.ctypes.data
        H = cat.n_halos
        N = int(pid32.shape[0]) if pid32 is not None else cat.n_tuples
        c.n_halos, c.n_tuples = H, N
        c.halo_id = arr(cat.halo_id, np.uint64)
        c.npart = arr(cat.npart, np.int32, (H, T))
        c.n_orphan_steps = arr(cat.n_orphan_steps, np.int32) if cat.n_orphan_steps is not None else None
        c.is_token = arr(cat.is_token, np.uint8) if cat.is_token is not None else None
        if device_ptrs is not None:
            c.pid, c.halo, c.type = device_ptrs
            self._check(self.lib.metro_catalogue_adopt_device(self.h, slot, C.byref(prm), C.byref(c)))
        elif pid32 is not None:
            assert halo_offset is not None
            c.pid, c.halo = None, None
            c.type = arr(cat.type, np.uint8) if cat.type is not None else None
            off = arr(halo_offset, np.int64, (H + 1,))
            self._check(self.lib.metro_catalogue_upload_csr32(self.h, slot, C.byref(prm), C.byref(c), arr(pid32, np.uint32), pid_base, off))
        else:
            c.pid = arr(cat.pid, np.uint64)
            c.type = arr(cat.type, np.uint8) if cat.type is not None else None
            if halo_offset is not None:
                c.halo = None
                off = arr(halo_offset, np.int64, (H + 1,))
                self._check(self.lib.metro_catalogue_upload_csr(self.h, slot, C.byref(prm), C.byref(c), off))
            else:
                c.halo = arr(cat.halo, np.uint32)
                self._check(self.lib.metro_catalogue_upload(self.h, slot, C.byref(prm), C.byref(c)))
        self.nptypes = T

    def catalogue_sizes(self, slot: int):
        nh, nt = C.c_int64(), C.c_int64()
        self._check(self.lib.metro_catalogue_info(self.h, slot, C.byref(nh), C.byref(nt)))
        return nh.value, nt.value

    def fetch_catalogue(self, slot: int, tuples: bool = True) -> HostCatalogue:
        H, N = self.catalogue_sizes(slot)
        T = self.nptypes
        hid, npart = np.zeros(H, np.uint64), np.zeros((H, T), np.int32)
        orph, tok = np.zeros(H, np.int32), np.zeros(H, np.uint8)
        self._check(self.lib.metro_catalogue_fetch_halos(self.h, slot, _ptr(hid), _ptr(npart), _ptr(orph), _ptr(tok)))
        pid = halo = typ = None
        if tuples:
            pid, halo, typ = np.zeros(N, np.uint64), np.zeros(N, np.uint32), np.zeros(N, np.uint8)
            self._check(self.lib.metro_catalogue_fetch_tuples(self.h, slot, _ptr(pid), _ptr(halo), _ptr(typ)))
        return HostCatalogue(hid, npart, orph, tok, pid, halo, typ)

    def set_npart(self, slot: int, npart: np.ndarray):
        a = np.ascontiguousarray(npart, np.int32)
        self._check(self.lib.metro_catalogue_set_npart(self.h, slot, _ptr(a)))

    # -- FindProgenitors(0,1) + FindProgenitors(1,0) + CleanTrees -----------------------------
    def match_pair(self, prm: Params) -> dict:
        info = _lib.ResultInfo()
        self._check(self.lib.metro_match_pair(self.h, C.byref(prm), C.byref(info)))
        return {k: (list(getattr(info, k)) if k in ("n_halos", "raw_count", "n_tuples") else getattr(info, k))
                for k, _ in _lib.ResultInfo._fields_}

    def fetch_result(self, info: dict, raw: bool = True, alloc=None) -> MatchResult:
        """alloc(shape, dtype) -> ndarray: where the results land (default: fresh numpy arrays; pass an allocator of pinned memory to
        receive them at full PCIe rate)."""
        T = self.nptypes
        H = info["n_halos"]
        M = info["raw_count"]
        new = alloc or (lambda shape, dtype: np.zeros(shape, dtype))
        ra = _lib.ResultArrays()
        raw_off = [new(H[d] + 1 if raw else 0, np.int64) for d in range(2)]
        raw_prog = [new(M[d] if raw else 0, np.int32) for d in range(2)]
        raw_nc = [new((M[d] if raw else 0, T), np.int32) for d in range(2)]
        raw_merit = [new(M[d] if raw else 0, np.uint32) for d in range(2)]
        if raw:
            for d in range(2):
                ra.raw_off[d], ra.raw_prog[d] = _ptr(raw_off[d]), _ptr(raw_prog[d])
                ra.raw_ncommon[d], ra.raw_merit[d] = _ptr(raw_nc[d]), _ptr(raw_merit[d])
        nt, cc, nk = info["n_trees"], info["clean_count"], info["n_tokens"]
        tree_main, tree_orphan = new(nt, np.int32), new(nt, np.uint8)
        tree_off = new(nt + 1, np.int64)
        clean_prog, clean_nc = new(cc, np.int32), new((cc, T), np.int32)
        token = new(nk, np.int32)
        ra.tree_main, ra.tree_orphan, ra.tree_off = _ptr(tree_main), _ptr(tree_orphan), _ptr(tree_off)
        ra.clean_prog, ra.clean_ncommon, ra.token_halo = _ptr(clean_prog), _ptr(clean_nc), _ptr(token)
        self._check(self.lib.metro_result_fetch(self.h, C.byref(ra)))
        return MatchResult(info, raw_off, raw_prog, raw_nc, raw_merit, tree_main, tree_orphan, tree_off,
                           clean_prog, clean_nc, token)

    def find_progenitors_and_clean(self, prm: Params, raw: bool = True) -> MatchResult:
        return self.fetch_result(self.match_pair(prm), raw=raw)

    def timings(self) -> dict:
        t = _lib.Timings()
        self._check(self.lib.metro_get_timings(self.h, C.byref(t)))
        d = {k: getattr(t, k) for k in ("total", "pack_hist", "sort", "join", "select", "exchange", "sort_launches", "total_launches", "path", "resident", "exchange_sendrecv")}
        d["sort_pass_ms"] = list(t.sort_pass_ms)
        return d

    # -- ShiftHalosPartsGrids ------------------------------------------------------------------
    def shift_halos_parts(self, prm: Params):
        self._check(self.lib.metro_shift(self.h, C.byref(prm)))

    # -- multi-GPU ---------------------------------------------------------------------------------
    def comm_unique_id(self) -> bytes:
        buf = C.create_string_buffer(128)
        self._check(self.lib.metro_comm_unique_id(buf))
        return buf.raw

    def comm_init(self, rank: int, world: int, unique_id: bytes):
        self._check(self.lib.metro_comm_init(self.h, rank, world, C.c_char_p(unique_id)))

    # -- building blocks -------------------------------------------------------------------------
    def merit_bits(self, n_comm, n_ph0, n_ph1, i_m) -> np.ndarray:
        a = [np.ascontiguousarray(x, np.int32) for x in (n_comm, n_ph0, n_ph1, i_m)]
        out = np.zeros(a[0].shape[0], np.uint32)
        self._check(self.lib.metro_merit_bits(self.h, out.shape[0], *[_ptr(x) for x in a], _ptr(out)))
        return out

    def sort_u64_device(self, keys_in, keys_out, vals_in, vals_out, n, begin_bit, end_bit):
        """Raw device pointers (ints)."""
        self._check(self.lib.metro_sort_u64(self.h, keys_in, keys_out, vals_in, vals_out, n, begin_bit, end_bit))

    def synth_pair_device(self, prm: Params, spec: SynthSpec):
        self._check(self.lib.metro_synth_pair_device(self.h, C.byref(prm), C.byref(spec)))
        self.nptypes = prm.nptypes


def synth_pair_host(prm: Params, spec: SynthSpec, slot: int) -> HostCatalogue:
    lib = _lib.load()
    nh, nt = C.c_int64(), C.c_int64()
    rc = lib.metro_synth_pair_host(C.byref(prm), C.byref(spec), slot, C.byref(nh), C.byref(nt), None, None, None, None, None)
    if rc:
        raise MetroError(rc, lib.metro_last_error(None).decode())
    H, N, T = nh.value, nt.value, prm.nptypes
    hid, npart = np.zeros(H, np.uint64), np.zeros((H, T), np.int32)
    pid, halo, typ = np.zeros(N, np.uint64), np.zeros(N, np.uint32), np.zeros(N, np.uint8)
    rc = lib.metro_synth_pair_host(C.byref(prm), C.byref(spec), slot, C.byref(nh), C.byref(nt), _ptr(hid), _ptr(npart),
                                   _ptr(pid), _ptr(halo), _ptr(typ))
    if rc:
        raise MetroError(rc, lib.metro_last_error(None).decode())
    return HostCatalogue(hid, npart, np.zeros(H, np.int32), np.zeros(H, np.uint8), pid, halo, typ)
