"""metrocpp_b200 -- B200-native progenitor matching (MetroCPP's FindProgenitors / CleanTrees path).

Only the path's pieces live here: csrc/ (CUDA kernels + the C ABI of include/metro_b200.h), the
ctypes loader and a thin Python mirror of the reference's call sequence.  Importing the engine
without the built shared library raises; there is no CPU fallback.
"""
from ._lib import FLAVOUR_GATHER, FLAVOUR_ZOOM_DUP, LIB_PATH, Params, SynthSpec  # noqa: F401
from .engine import HostCatalogue, MatchResult, MetroEngine, MetroError, make_params, synth_pair_host  # noqa: F401
