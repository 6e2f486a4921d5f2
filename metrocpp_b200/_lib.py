"""ctypes loader for libmetro_b200.so (the C ABI declared in include/metro_b200.h).

The library is built in-tree by `make -C metrocpp_b200/csrc` (see __graft_entry__.build()).
There is no fallback: if the shared object is missing the import of the engine fails loudly.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("METRO_B200_LIB") or os.path.join(HERE, "libmetro_b200.so")   # override: tuning variants only

OK, ERR_INVALID, ERR_CUDA, ERR_UNSUPPORTED, ERR_STATE, ERR_COMM = range(6)
FLAVOUR_GATHER, FLAVOUR_ZOOM_DUP = 0, 1


class Params(C.Structure):
    _fields_ = [("nptypes", C.c_int32), ("min_part_cmp", C.c_int32), ("min_part_halo", C.c_int32),
                ("fac_orphan_steps", C.c_int32), ("max_orphan_steps", C.c_int32), ("flavour", C.c_int32)]


class Catalogue(C.Structure):
    _fields_ = [("n_halos", C.c_int64), ("halo_id", C.c_void_p), ("npart", C.c_void_p),
                ("n_orphan_steps", C.c_void_p), ("is_token", C.c_void_p), ("n_tuples", C.c_int64),
                ("pid", C.c_void_p), ("halo", C.c_void_p), ("type", C.c_void_p)]


class ResultInfo(C.Structure):
    _fields_ = [("n_halos", C.c_int64 * 2), ("raw_count", C.c_int64 * 2), ("n_trees", C.c_int64),
                ("clean_count", C.c_int64), ("n_tokens", C.c_int64), ("n_untracked", C.c_int64),
                ("n_hits", C.c_int64), ("n_pairs", C.c_int64), ("n_tuples", C.c_int64 * 2),
                ("pid_bits", C.c_int32), ("halo_bits", C.c_int32), ("type_bits", C.c_int32),
                ("sort_passes", C.c_int32), ("n_table_hits", C.c_int64), ("exchange_bytes", C.c_int64),
                ("owned_lo", C.c_int64), ("owned_hi", C.c_int64)]


class ResultArrays(C.Structure):
    _fields_ = [("raw_off", C.c_void_p * 2), ("raw_prog", C.c_void_p * 2), ("raw_ncommon", C.c_void_p * 2),
                ("raw_merit", C.c_void_p * 2), ("tree_main", C.c_void_p), ("tree_orphan", C.c_void_p),
                ("tree_off", C.c_void_p), ("clean_prog", C.c_void_p), ("clean_ncommon", C.c_void_p),
                ("token_halo", C.c_void_p)]


class Timings(C.Structure):
    _fields_ = [("total", C.c_float), ("pack_hist", C.c_float), ("sort", C.c_float), ("join", C.c_float),
                ("select", C.c_float), ("exchange", C.c_float), ("sort_launches", C.c_int32), ("total_launches", C.c_int32),
                ("sort_pass_ms", C.c_float * 8), ("path", C.c_int32), ("resident", C.c_int32), ("exchange_sendrecv", C.c_float)]


class SynthSpec(C.Structure):
    _fields_ = [("seed", C.c_uint64), ("n_halos", C.c_int64), ("n_tuples", C.c_int64), ("pid_bits", C.c_int32),
                ("shard", C.c_int32), ("n_shards", C.c_int32)]


# every symbol include/metro_b200.h declares: name -> (restype, argtypes)
_P = C.c_void_p
SYMBOLS = {
    "metro_ctx_create": (C.c_int, [C.c_int, C.POINTER(_P)]),
    "metro_ctx_destroy": (None, [_P]),
    "metro_last_error": (C.c_char_p, [_P]),
    "metro_version": (C.c_char_p, []),
    "metro_ctx_set_stream": (C.c_int, [_P, _P]),
    "metro_catalogue_upload": (C.c_int, [_P, C.c_int, C.POINTER(Params), C.POINTER(Catalogue)]),
    "metro_catalogue_upload_csr": (C.c_int, [_P, C.c_int, C.POINTER(Params), C.POINTER(Catalogue), _P]),
    "metro_catalogue_upload_csr32": (C.c_int, [_P, C.c_int, C.POINTER(Params), C.POINTER(Catalogue), _P, C.c_uint64, _P]),
    "metro_catalogue_adopt_device": (C.c_int, [_P, C.c_int, C.POINTER(Params), C.POINTER(Catalogue)]),
    "metro_match_pair": (C.c_int, [_P, C.POINTER(Params), C.POINTER(ResultInfo)]),
    "metro_result_fetch": (C.c_int, [_P, C.POINTER(ResultArrays)]),
    "metro_get_timings": (C.c_int, [_P, C.POINTER(Timings)]),
    "metro_shift": (C.c_int, [_P, C.POINTER(Params)]),
    "metro_catalogue_info": (C.c_int, [_P, C.c_int, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    "metro_catalogue_fetch_halos": (C.c_int, [_P, C.c_int, _P, _P, _P, _P]),
    "metro_catalogue_fetch_tuples": (C.c_int, [_P, C.c_int, _P, _P, _P]),
    "metro_catalogue_set_npart": (C.c_int, [_P, C.c_int, _P]),
    "metro_comm_unique_id": (C.c_int, [_P]),
    "metro_comm_init": (C.c_int, [_P, C.c_int, C.c_int, _P]),
    "metro_sort_u64": (C.c_int, [_P, _P, _P, _P, _P, C.c_int64, C.c_int, C.c_int]),
    "metro_merit_bits": (C.c_int, [_P, C.c_int64, _P, _P, _P, _P, _P]),
    "metro_synth_pair_device": (C.c_int, [_P, C.POINTER(Params), C.POINTER(SynthSpec)]),
    "metro_synth_pair_host": (C.c_int, [C.POINTER(Params), C.POINTER(SynthSpec), C.c_int, C.POINTER(C.c_int64),
                                        C.POINTER(C.c_int64), _P, _P, _P, _P, _P]),
}

_lib = None


def load() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.isfile(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(or `make -C metrocpp_b200/csrc`). metrocpp_b200 has no CPU fallback.")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(lib, name)      # AttributeError if the .so does not export it
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib
