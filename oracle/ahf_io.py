"""AHF text I/O used by the parity tests.  TEST INFRASTRUCTURE ONLY.

Writes the two AHF catalogue files the reference ingests and parses the
``.mtree`` text it emits.  Grammar follows the reference readers/writer:

* ``_halos``     : IOSettings::ReadHalos / ReadLineAHF (reference
                   src/IOSettings.cpp:712-811, 914-996): ``#`` lines skipped, then
                   whitespace columns ID(1) hostHalo(2) numSubStruct(3) Mvir(4)
                   npart(5) Xc Yc Zc VXc VYc VZc ... (44 columns scanned).
* ``_particles`` : IOSettings::ReadParticles (src/IOSettings.cpp:551-708): line 1 =
                   number of halos in the file; per halo ``nPart haloID`` followed
                   by nPart lines ``particleID [type]``.
* ``.mtree``     : IOSettings::WriteTree (src/IOSettings.cpp:1074-1138).
"""
from __future__ import annotations

import os
from dataclasses import dataclass, field

import numpy as np


@dataclass
class Snapshot:
    """One AHF catalogue in memory: halos in file order + CSR of (pid, type) tuples."""

    number: int                 # snapshot number (file name _NNN)
    redshift: float
    halo_id: np.ndarray         # uint64[H]
    npart: np.ndarray           # int32[H]  catalogue column 5
    offsets: np.ndarray         # int64[H+1] CSR into pid / ptype
    pid: np.ndarray             # uint64[N]
    ptype: np.ndarray           # uint8[N]
    pos: np.ndarray = field(default=None)   # float32[H,3]
    vel: np.ndarray = field(default=None)   # float32[H,3]

    @property
    def n_halos(self) -> int:
        return int(self.halo_id.shape[0])

    @property
    def n_tuples(self) -> int:
        return int(self.pid.shape[0])

    def halo_index_per_tuple(self) -> np.ndarray:
        counts = np.diff(self.offsets)
        return np.repeat(np.arange(self.n_halos, dtype=np.uint32), counts)


def file_stem(prefix: str, number: int, redshift: float, chunk: int | None) -> str:
    """File naming of IOSettings::DistributeFilesAmongTasks (src/IOSettings.cpp:447-547)."""
    if chunk is None:
        return f"{prefix}{number:03d}.z{redshift:.3f}"
    return f"{prefix}{number:03d}.{chunk:04d}.z{redshift:.3f}"


def write_snapshot(directory: str, snap: Snapshot, prefix: str = "snapshot_",
                   chunk: int | None = None, with_types: bool = True) -> tuple[str, str]:
    """Write <stem>.AHF_halos and <stem>.AHF_particles; returns both paths."""
    os.makedirs(directory, exist_ok=True)
    stem = os.path.join(directory, file_stem(prefix, snap.number, snap.redshift, chunk))
    H = snap.n_halos
    pos = snap.pos if snap.pos is not None else np.full((H, 3), 50000.0, np.float32)
    vel = snap.vel if snap.vel is not None else np.zeros((H, 3), np.float32)
    halo_path = stem + ".AHF_halos"
    with open(halo_path, "w") as f:
        f.write("#ID(1) hostHalo(2) numSubStruct(3) Mvir(4) npart(5) Xc(6) Yc(7) Zc(8) "
                "VXc(9) VYc(10) VZc(11) Rvir(12)\n")
        for h in range(H):
            f.write("%d 0 0 %.6e %d %.3f %.3f %.3f %.2f %.2f %.2f 100.0 50.0 10.0 0.0 0.0 150.0\n" % (
                int(snap.halo_id[h]), 1.0e8 * float(snap.npart[h]), int(snap.npart[h]),
                pos[h, 0], pos[h, 1], pos[h, 2], vel[h, 0], vel[h, 1], vel[h, 2]))
    part_path = stem + ".AHF_particles"
    # vectorised body: one text line per tuple, halo header lines interleaved
    lines = []
    pid = snap.pid
    pt = snap.ptype
    off = snap.offsets
    with open(part_path, "w") as f:
        f.write(f"{H}\n")
        for h in range(H):
            a, b = int(off[h]), int(off[h + 1])
            f.write(f"{b - a} {int(snap.halo_id[h])}\n")
            if b > a:
                if with_types:
                    body = "\n".join(f"{p} {t}" for p, t in zip(pid[a:b].tolist(), pt[a:b].tolist()))
                else:
                    body = "\n".join(map(str, pid[a:b].tolist()))
                f.write(body)
                f.write("\n")
    del lines
    return halo_path, part_path


@dataclass
class MTreeEntry:
    main_id: int
    main_npart: int
    orphan: int
    progs: list  # of (npart_prog, prog_id, n_common)


def parse_mtree(text: str) -> list[MTreeEntry]:
    """Parse the text written by IOSettings::WriteTree (src/IOSettings.cpp:1113-1132)."""
    out: list[MTreeEntry] = []
    rows = [ln.split() for ln in text.splitlines() if ln and not ln.startswith("#")]
    i = 0
    while i < len(rows):
        main_id, main_np, nprog, orphan = (int(x) for x in rows[i])
        progs = []
        for j in range(nprog):
            a, b, c = (int(x) for x in rows[i + 1 + j])
            progs.append((a, b, c))
        out.append(MTreeEntry(main_id, main_np, orphan, progs))
        i += 1 + nprog
    return out


MTREE_HEADER = ("# ID host(1)   N particles host(2)   Num. progenitors(3)  Orphan[0=no, 1=yes](4)\n"
                "# Total particles (1)   ID progenitor(2)   Particles in common (3)\n")
