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


def _parse_only(cfg, threads):
    r = subprocess.run([BIN, "--parse-only", cfg], capture_output=True, text=True, env=dict(os.environ, METRO_INGEST_THREADS=str(threads)))
    assert r.returncode == 0, r.stdout + r.stderr
    return [dict(f.split("=") for f in ln.split()[1:]) for ln in r.stdout.splitlines() if ln.startswith("PARSED")], r.stderr


def test_ingest_grammar_edge_cases(tmp_path):
    """_particles grammar corners through the chunked parser (reference src/IOSettings.cpp:599-689): several blocks in one
    file, blank lines between blocks, halos listed in a different order than the _halos file, a halo ID that is not in
    the catalogue (dropped with a warning), a halo without particles, CRLF-free trailing spaces."""
    inp = tmp_path / "input"
    inp.mkdir(); (tmp_path / "output").mkdir()
    halos = {101: [5, 6, 7], 102: [], 103: [8, 9], 104: [10, 11, 12, 13]}
    for number, z in ((10, "0.000"), (9, "0.100")):
        with open(inp / f"snapshot_{number:03d}.z{z}.AHF_halos", "w") as f:
            f.write("#ID(1) hostHalo(2) numSubStruct(3) Mvir(4) npart(5) Xc Yc Zc VXc VYc VZc\n")
            for hid, parts in halos.items():
                f.write(f"{hid} 0 0 1.0e10 {len(parts)} 1 2 3 4 5 6\n")
        with open(inp / f"snapshot_{number:03d}.z{z}.AHF_particles", "w") as f:
            # block 1: halos 103 and 101 (out of catalogue order), then a blank line, block 2: 999 (unknown), 104, 102 (empty)
            f.write("2\n2 103\n8 1\n9 1  \n3 101\n5 1\n6 1\n7 1\n\n3\n2 999\n77 1\n78 1\n4 104\n10 1\n11 1\n12 1\n13 1\n0 102\n")
    cfg = tmp_path / "config.cfg"
    cfg.write_text("inputFormat = AHF\nhaloSuffix = AHF_halos\npartSuffix = AHF_particles\nhaloPrefix = snapshot_\npartPrefix = snapshot_\n"
                   f"cpuString = .\nsplitString = _\npathMetroCpp = {tmp_path}/\npathInput = {inp}/\nboxSize = 1000.0\nnChunks = 1\nnSnapsUse = 2\n"
                   f"nGrid = 4\npathOutput = {tmp_path}/output/\noutPrefix = t_\noutSuffix = mtree\nminPartHalo = 1\nminPartCmp = 1\nfacOrphanSteps = 15\n"
                   "cosmologicalModel = Planck\nnTreeChunks = 1\ntreeFlavour = zoom\nnPTypes = 2\nnoPType = 1\n")
    # expected tuples in file order: (pid, halo index in the _halos order 101,102,103,104 -> 0,1,2,3)
    exp = [(8, 2), (9, 2), (5, 0), (6, 0), (7, 0), (10, 3), (11, 3), (12, 3), (13, 3)]
    exp_ck = fnv(p + 0x9E3779B97F4A7C15 * (h + 1) + 1 for p, h in exp)
    for threads in (1, 2, 5, 16):
        parsed, err = _parse_only(str(cfg), threads)
        assert len(parsed) == 2
        for kv in parsed:
            assert int(kv["halos"]) == 4 and int(kv["tuples"]) == len(exp) and int(kv["tuple_ck"]) == exp_ck, (threads, kv)
        assert "2 particle lines belong to halo IDs that are not in the halo catalogue" in err


def test_ingest_fast_tokeniser_lexical_corners(tmp_path):
    """The eight-digits-at-a-time tokeniser of the _particles reader against a plain Python parse of the same text: 1-19 digit IDs
    (8-, 9-, 16- and 17-digit ones included), leading / trailing blanks and tabs, CRLF line ends, extra tokens behind the type, a
    last line without a newline; whole file in one chunk and cut into 3 / 8 / 61 chunks (every chunk ends in a bounds-checked
    tail of 64 bytes, so the cuts move the hand-over between the two code paths across all of these line shapes)."""
    import random
    rnd = random.Random(20261020)
    inp = tmp_path / "input"
    inp.mkdir(); (tmp_path / "output").mkdir()
    n_halos = 300
    sizes = [rnd.randint(0, 120) for _ in range(n_halos)]
    digits = [1, 2, 7, 8, 9, 10, 15, 16, 17, 18, 19]
    lines, exp = [f"{n_halos}"], []
    for h, n in enumerate(sizes):
        lines.append(f"{n} {1000 + h}" + rnd.choice(["", " ", "\r", "  \t"]))
        for _ in range(n):
            nd = rnd.choice(digits)
            pid = rnd.randint(10 ** (nd - 1) if nd > 1 else 0, min(10 ** nd - 1, 2 ** 63 - 1))
            ty = rnd.choice([0, 1, 4, 5])
            lead = rnd.choice(["", "", "", " ", "\t ", "   "])
            mid = rnd.choice([" ", " ", "\t", "    "])
            tail = rnd.choice(["", "", "", " ", "\r", " \r", " 77 extra", "\tx"])
            lines.append(f"{lead}{pid}{mid}{ty}{tail}")
            exp.append((pid, h, ty))
    text = "\n".join(lines)                                   # no newline behind the last line
    for number, z in ((10, "0.000"), (9, "0.100")):
        with open(inp / f"snapshot_{number:03d}.z{z}.AHF_halos", "w") as f:
            f.write("#ID(1) hostHalo(2) numSubStruct(3) Mvir(4) npart(5) Xc Yc Zc VXc VYc VZc\n")
            for h, n in enumerate(sizes):
                f.write(f"{1000 + h} 0 0 1.0e10 {n} 1 2 3 4 5 6\n")
        with open(inp / f"snapshot_{number:03d}.z{z}.AHF_particles", "w", newline="") as f:
            f.write(text)
    cfg = tmp_path / "config.cfg"
    cfg.write_text("inputFormat = AHF\nhaloSuffix = AHF_halos\npartSuffix = AHF_particles\nhaloPrefix = snapshot_\npartPrefix = snapshot_\n"
                   f"cpuString = .\nsplitString = _\npathMetroCpp = {tmp_path}/\npathInput = {inp}/\nboxSize = 1000.0\nnChunks = 1\nnSnapsUse = 2\n"
                   f"nGrid = 4\npathOutput = {tmp_path}/output/\noutPrefix = t_\noutSuffix = mtree\nminPartHalo = 1\nminPartCmp = 1\nfacOrphanSteps = 15\n"
                   "cosmologicalModel = Planck\nnTreeChunks = 1\ntreeFlavour = zoom\nnPTypes = 6\nnoPType = 0\n")
    exp_ck = fnv(p + 0x9E3779B97F4A7C15 * (h + 1) + t for p, h, t in exp)
    assert len(text) > 100_000
    for threads in (1, 3, 8, 61):
        parsed, _ = _parse_only(str(cfg), threads)
        assert len(parsed) == 2
        for kv in parsed:
            assert int(kv["halos"]) == n_halos and int(kv["tuples"]) == len(exp) and int(kv["tuple_ck"]) == exp_ck, (threads, kv)


# ---- .mtree reader compatibility ---------------------------------------------------------------------------------
def reader_rules(text):
    """What the reference's own post-processing reader makes of a .mtree file (python/mtreelib/read_mtree.py:98-145): a line of 4
    fields opens a tree (descendant ID, its particles, number of progenitors, orphan flag), the FIRST following line of 3 fields is its
    main progenitor (particles, ID, common particles); everything else is skipped.  -> {descID: (progID, common, descNPart, descNProg)}"""
    out, n_prog_line, desc = {}, 0, None
    for line in text.splitlines():
        f = line.split()
        if len(f) == 4:
            desc, n_prog_line = (f[0], int(f[1]), int(f[2])), 0
        if len(f) == 3:
            n_prog_line += 1
            if n_prog_line == 1:
                out[desc[0]] = (f[1], int(f[2]), desc[1], desc[2])
    return out


@pytest.mark.skipif(not os.path.isfile("/root/reference/python/mtreelib/read_mtree.py"), reason="needs the reference's python/ tree")
@pytest.mark.parametrize("name", ["chain6_tokens", "fullbox_chain20"])
def test_reference_python_reader_parses_our_mtree_layout(tmp_path, name):
    """The reference's reader, imported from /root/reference/python, on the .mtree text of the fixtures (byte-identical to what
    IOSettings::WriteTree of the host program writes, see the GPU test below): same main-branch dictionary as the restated rules,
    and the file names are the ones ReadSettings.gen_file_list builds (baseFile + '%03d.' + '%d.' + suffFile)."""
    import sys
    sys.path.insert(0, "/root/reference/python")
    try:
        from mtreelib import read_mtree
    finally:
        sys.path.pop(0)
    snaps, params, mtree = golden_io.load(name)
    flavour = "gather" if "gather" in mtree else sorted(mtree)[0]
    numbers = sorted(mtree[flavour], reverse=True)
    for number in numbers:
        path = tmp_path / f"test_{number:03d}.0.mtree"
        path.write_text(mtree[flavour][number])
        got = read_mtree.read_metrocpp_tree(str(path))
        exp = reader_rules(mtree[flavour][number])
        assert set(got) == set(exp) and len(exp) > 0
        for k, dp in got.items():
            assert (dp.progID, int(dp.progNPart), int(dp.descNPart), int(dp.descNProg)) == exp[k]
    settings = read_mtree.ReadSettings(str(tmp_path / "test_"), "mtree", 1, numbers[0], len(numbers))
    assert [f[0] for f in settings.treeFiles] == [str(tmp_path / f"test_{n:03d}.0.mtree") for n in numbers]


@pytest.mark.gpu
def test_host_program_output_reads_like_the_reference_output(tmp_path):
    snaps, params, mtree = golden_io.load("chain6_tokens")
    cfg, out = write_case(str(tmp_path), snaps, params, "zoom")
    r = subprocess.run([BIN, cfg], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    for number, text in mtree["zoom"].items():
        got = open(os.path.join(out, f"test_{number:03d}.0.mtree")).read()       # the name the reference's reader looks for
        assert reader_rules(got) == reader_rules(text) and len(reader_rules(got)) > 0
