"""Drive the compiled, unmodified reference (oracle/_ref/MetroCPP_*).  TEST INFRASTRUCTURE ONLY.

Only usable where both ``oracle/_ref`` binaries and ``/root/reference`` exist
(the build container): the reference shells out to ``scripts/find_*.sh`` under
``pathMetroCpp`` (reference src/IOSettings.cpp:237-376), so a scratch
``pathMetroCpp`` is assembled from symlinks into /root/reference plus a fresh
writable ``tmp/``.  Used to (a) validate the C restatement and (b) generate the
golden ``.mtree`` fixtures under tests/golden/ (see tests/golden/make_golden.py).
"""
from __future__ import annotations

import os
import shutil
import subprocess
import tempfile
import time

from . import ahf_io

HERE = os.path.dirname(os.path.abspath(__file__))
REF_ROOT = os.environ.get("METRO_REFERENCE_ROOT", "/root/reference")
FLAVOURS = {"zoom": "MetroCPP_zoom", "gather": "MetroCPP_gather", "zoom_t6": "MetroCPP_zoom_t6"}


def binary(flavour: str) -> str:
    return os.path.join(HERE, "_ref", FLAVOURS[flavour])


def available(flavour: str = "zoom") -> bool:
    return os.path.isfile(binary(flavour)) and os.path.isdir(os.path.join(REF_ROOT, "scripts"))


def write_config(path: str, *, metro: str, inp: str, out: str, n_snaps: int, fullbox: bool,
                 params: dict) -> None:
    lines = [
        "inputFormat = AHF",
        "haloSuffix = AHF_halos",
        "partSuffix = AHF_particles",
        "haloPrefix = snapshot_",
        "partPrefix = snapshot_",
        "cpuString = " + (".0000." if fullbox else "."),
        "splitString = _",
        f"pathMetroCpp = {metro}/",
        f"pathInput = {inp}/",
        "boxSize = 100000.0",
        "nChunks = 1",
        f"nSnapsUse = {n_snaps}",
        "nGrid = 4",
        f"pathOutput = {out}/",
        "outPrefix = test_",
        "outSuffix = mtree",
        f"minPartHalo = {params.get('minPartHalo', 30)}",
        f"minPartCmp = {params.get('minPartCmp', 10)}",
        f"facOrphanSteps = {params.get('facOrphanSteps', 15)}",
    ]
    if params.get("maxOrphanSteps") is not None:
        lines.append(f"maxOrphanSteps = {params['maxOrphanSteps']}")
    lines += ["cosmologicalModel = Planck", "nTreeChunks = 1"]
    with open(path, "w") as f:
        f.write("\n".join(lines) + "\n")


def run_reference(snapshots: list[ahf_io.Snapshot], params: dict, flavour: str = "zoom",
                  keep: str | None = None, timeout: float = 3600.0) -> dict:
    """Run the reference over a chain (any order; sorted by number) and return
    ``{"mtree": {later_snapshot_number: text}, "timing": {...}, "stdout": str}``."""
    if not available(flavour):
        raise RuntimeError(f"reference flavour {flavour!r} is not built / /root/reference missing")
    fullbox = flavour == "gather"
    with_types = flavour == "zoom_t6"
    work = keep or tempfile.mkdtemp(prefix="metro_ref_")
    metro, inp, out = (os.path.join(work, d) for d in ("metro", "input", "output"))
    for d in (metro, inp, out, os.path.join(metro, "tmp")):
        os.makedirs(d, exist_ok=True)
    for d in ("scripts", "data"):
        link = os.path.join(metro, d)
        if not os.path.exists(link):
            os.symlink(os.path.join(REF_ROOT, d), link)
    for f in os.listdir(os.path.join(metro, "tmp")):
        os.remove(os.path.join(metro, "tmp", f))          # stale-cache hazard
    for s in snapshots:
        ahf_io.write_snapshot(inp, s, chunk=0 if fullbox else None, with_types=with_types)
    cfg = os.path.join(work, "config.cfg")
    write_config(cfg, metro=metro, inp=inp, out=out, n_snaps=len(snapshots), fullbox=fullbox,
                 params=params)
    t0 = time.time()
    proc = subprocess.run([binary(flavour), cfg], cwd=work, capture_output=True, text=True,
                          timeout=timeout)
    wall = time.time() - t0
    mtree = {}
    for fn in sorted(os.listdir(out)):
        if fn.endswith(".mtree"):
            num = int(fn[len("test_"):len("test_") + 3])
            with open(os.path.join(out, fn)) as f:
                mtree[num] = f.read()
    timing = {}
    for fn in os.listdir(out):
        if fn.startswith("timing_log"):
            with open(os.path.join(out, fn)) as f:
                timing["log"] = f.read()
    result = {"mtree": mtree, "timing": timing, "stdout": proc.stdout, "stderr": proc.stderr,
              "wall_s": wall, "returncode": proc.returncode, "work": work}
    if keep is None:
        shutil.rmtree(work, ignore_errors=True)
    return result
