"""`mpirun -n K <binary> <args>` for the multi-process MPI shim (oracle/mpi_shim_mp).  TEST INFRASTRUCTURE ONLY.

Creates the zeroed control segment in /dev/shm, starts K ranks of the binary with MPISHIM_RANK / MPISHIM_SIZE / MPISHIM_TAG set,
waits for them and removes every /dev/shm object of the run.

    python -m oracle.mpirun_shim -n 4 oracle/_ref/MetroCPP_gather_mp config.cfg
"""
from __future__ import annotations

import glob
import os
import subprocess
import sys
import time

BOX_BYTES = 64          # sizeof(Box) in mpi_mp.cpp


def run(n: int, argv: list[str], cwd: str | None = None, timeout: float | None = None, capture: bool = True) -> list[subprocess.CompletedProcess]:
    tag = f"metro_mpishim_{os.getpid()}_{int(time.time() * 1e3) % 10**9}"
    ctl = f"/dev/shm/{tag}.ctl"
    with open(ctl, "wb") as f:
        f.write(b"\0" * (BOX_BYTES * n * n))
    procs = []
    try:
        for r in range(n):
            env = dict(os.environ, MPISHIM_RANK=str(r), MPISHIM_SIZE=str(n), MPISHIM_TAG=tag)
            procs.append(subprocess.Popen(argv, cwd=cwd, env=env, stdout=subprocess.PIPE if capture else None,
                                          stderr=subprocess.STDOUT if capture else None, text=True))
        out = []
        deadline = None if timeout is None else time.time() + timeout
        for p in procs:
            left = None if deadline is None else max(1.0, deadline - time.time())
            so, _ = p.communicate(timeout=left)
            out.append(subprocess.CompletedProcess(argv, p.returncode, so, None))
        return out
    finally:
        for p in procs:
            if p.poll() is None:
                p.kill()
        for f in glob.glob(f"/dev/shm/{tag}.*"):
            try:
                os.remove(f)
            except OSError:
                pass


def main():
    if len(sys.argv) < 4 or sys.argv[1] != "-n":
        print(__doc__)
        return 2
    res = run(int(sys.argv[2]), sys.argv[3:], capture=False)
    return max(r.returncode for r in res)


if __name__ == "__main__":
    sys.exit(main())
