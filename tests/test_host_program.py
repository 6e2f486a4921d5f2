"""The C++ host program above the C ABI (metrocpp_b200/host: Halo / MergerTree / IOSettings / main).

CPU: the AHF readers and the in-process snapshot discovery (`--parse-only`, no GPU needed).
GPU: AHF text in -> .mtree text out, byte-identical to the reference's own output (golden fixtures).
"""
import os
import subprocess

import numpy as np
import pytest

from oracle import ahf_io
from tests import golden_io

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "metrocpp_b200", "MetroCPP-b200")


def write_case(tmp, snaps, params, flavour, fullbox_names=False):
    inp, out = os.path.join(tmp, "input"), os.path.join(tmp, "output")
    os.makedirs(inp); os.makedirs(out)
    typed = flavour == "zoom_t6"
    for s in snaps:
        ahf_io.write_snapshot(inp, s, chunk=0 if fullbox_names else None, with_types=typed)
    cfg = os.path.join(tmp, "config.cfg")
    lines = ["# test config", "inputFormat = AHF", "haloSuffix = AHF_halos", "partSuffix = AHF_particles",
             "haloPrefix = snapshot_", "partPrefix = snapshot_", "cpuString = .", "splitString = _",
             f"pathMetroCpp = {tmp}/", f"pathInput = {inp}/", "boxSize = 100000.0", "nChunks = 1",
             f"nSnapsUse = {len(snaps)}", "nGrid = 4", f"pathOutput = {out}/", "outPrefix = test_", "outSuffix = mtree",
             f"minPartHalo = {params.get('minPartHalo', 30)}", f"minPartCmp = {params.get('minPartCmp', 10)}",
             f"facOrphanSteps = {params.get('facOrphanSteps', 15)}", "cosmologicalModel = Planck", "nTreeChunks = 1",
             f"treeFlavour = {'gather' if flavour == 'gather' else 'zoom'}",
             f"nPTypes = {6 if typed else 2}", f"noPType = {0 if typed else 1}"]
    if params.get("maxOrphanSteps") is not None:
        lines.append(f"maxOrphanSteps = {params['maxOrphanSteps']}")
    open(cfg, "w").write("\n".join(lines) + "\n")
    return cfg, out


def fnv(values):
    ck = 1469598103934665603
    for v in values:
        ck = ((ck ^ (v & 0xFFFFFFFFFFFFFFFF)) * 1099511628211) & 0xFFFFFFFFFFFFFFFF
    return ck


@pytest.mark.parametrize("fullbox_names", [False, True])
def test_ahf_ingest_matches_python_reader(tmp_path, fullbox_names):
    snaps, params, _ = golden_io.load("multi_species")
    cfg, _ = write_case(str(tmp_path), snaps, params, "zoom_t6", fullbox_names)
    r = subprocess.run([BIN, "--parse-only", cfg], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    parsed = [ln for ln in r.stdout.splitlines() if ln.startswith("PARSED")]
    assert len(parsed) == len(snaps)
    for ln, s in zip(parsed, sorted(snaps, key=lambda x: -x.number)):      # index 0 = latest snapshot
        kv = dict(f.split("=") for f in ln.split()[1:])
        assert int(kv["snapshot"]) == s.number and int(kv["halos"]) == s.n_halos and int(kv["tuples"]) == s.n_tuples
        halo = s.halo_index_per_tuple()
        exp = fnv(int(p) + 0x9E3779B97F4A7C15 * (int(h) + 1) + int(t) for p, h, t in zip(s.pid, halo, s.ptype))
        assert int(kv["tuple_ck"]) == exp
        assert int(kv["halo_ck"]) == fnv(int(i) + int(n) for i, n in zip(s.halo_id, s.npart))


def test_parallel_ingest_equals_serial(tmp_path):
    """The chunked multi-thread _particles parser (one entry per line, serial header chain, parallel copy-out)
    gives the same tuples as one thread, whatever the number of chunks."""
    snaps, params, _ = golden_io.load("fullbox_chain20")
    cfg, _ = write_case(str(tmp_path), snaps[:3], params, "gather", True)
    outs = []
    for nt in ("1", "3", "7", "64"):
        r = subprocess.run([BIN, "--parse-only", cfg], capture_output=True, text=True, env=dict(os.environ, METRO_INGEST_THREADS=nt))
        assert r.returncode == 0, r.stdout + r.stderr
        outs.append([ln for ln in r.stdout.splitlines() if ln.startswith("PARSED")])
    assert len(outs[0]) == 3 and all(o == outs[0] for o in outs[1:])


def test_missing_input_fails_loudly(tmp_path):
    cfg = tmp_path / "c.cfg"
    cfg.write_text("inputFormat = AHF\nhaloSuffix = AHF_halos\npartSuffix = AHF_particles\nhaloPrefix = snapshot_\n"
                   f"partPrefix = snapshot_\npathInput = {tmp_path}/nowhere/\npathOutput = {tmp_path}/\n")
    r = subprocess.run([BIN, "--parse-only", str(cfg)], capture_output=True, text=True)
    assert r.returncode != 0 and "cannot be opened" in r.stderr


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["chain6_tokens", "merit_tie", "multi_species_tokens", "fullbox_chain20", "hashed_ids"])
def test_host_program_reproduces_reference_mtree(tmp_path, name):
    snaps, params, mtree = golden_io.load(name)
    for flavour, expect in mtree.items():
        tmp = tmp_path / flavour
        tmp.mkdir()
        cfg, out = write_case(str(tmp), snaps, params, flavour, fullbox_names=(flavour == "gather"))
        r = subprocess.run([BIN, cfg], capture_output=True, text=True)
        assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
        for number, text in expect.items():
            got = open(os.path.join(out, f"test_{number:03d}.0.mtree")).read()
            assert got == text, f"{name}/{flavour}/{number}"
