"""Load the committed golden fixtures (tests/golden/*.npz, made by tests/golden/make_golden.py)."""
from __future__ import annotations

import glob
import json
import os

import numpy as np

from oracle.ahf_io import Snapshot

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def names():
    return sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN_DIR, "*.npz")))


def load(name):
    """-> (snapshots ascending by number, params dict, {flavour: {number: mtree text}})"""
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    meta = json.loads(bytes(z["meta"]).decode())
    snaps = []
    for number, red in zip(z["numbers"], z["redshifts"]):
        k = f"s{int(number)}_"
        snaps.append(Snapshot(int(number), float(red), z[k + "halo_id"], z[k + "npart"], z[k + "offsets"],
                              z[k + "pid"], z[k + "ptype"]))
    mtree = {fl: {int(k): v for k, v in d.items()} for fl, d in meta["mtree"].items()}
    return snaps, meta["params"], mtree
