"""Host-side AHF ingest throughput of MetroCPP-b200 (no GPU needed: --parse-only).  Writes a synthetic pair as AHF text
(oracle/synth_np.py + oracle/ahf_io.py), runs the reader with all cores and with one, prints its phase timers.
usage: python scripts/ingest_bench.py [n_halos] [n_tuples]     (default 40000 2e7: two 172 MB _particles files)"""
import os
import shutil
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import synth_np  # noqa: E402
from tests.test_host_program import BIN, write_case  # noqa: E402


def main():
    n_halos = int(float(sys.argv[1])) if len(sys.argv) > 1 else 40_000
    n_tuples = int(float(sys.argv[2])) if len(sys.argv) > 2 else 20_000_000
    tmp = tempfile.mkdtemp(prefix="metro_ingest_", dir=os.environ.get("TMPDIR", "/tmp"))
    try:
        snaps = synth_np.synth_pair(7, n_halos, n_tuples)
        cfg, _ = write_case(tmp, snaps, {"minPartHalo": 30}, "zoom", False)
        for f in sorted(os.listdir(os.path.join(tmp, "input"))):
            print(f"{f}: {os.path.getsize(os.path.join(tmp, 'input', f)) / 1e6:.1f} MB")
        for threads in (str(os.cpu_count() or 1), "1"):
            t0 = time.perf_counter()
            r = subprocess.run([BIN, "--parse-only", cfg], capture_output=True, text=True, env=dict(os.environ, METRO_INGEST_THREADS=threads, METRO_DEBUG="1"))
            assert r.returncode == 0, r.stderr
            print(f"threads {threads}: wall {time.perf_counter() - t0:.3f} s (two snapshots, including the checksums --parse-only prints)")
            print("\n".join(ln for ln in r.stderr.splitlines() if "ingest" in ln))
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


if __name__ == "__main__":
    main()
