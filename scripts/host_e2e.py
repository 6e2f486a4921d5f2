"""End-to-end comparison at SURVEY config-1 scale (zoom-style pair: 2e4 halos, 1e7 bound IDs per snapshot):
the unmodified compiled reference and the host program MetroCPP-b200 read the SAME AHF text files and must
write byte-identical .mtree files; prints the wall time of both and the phases of ours.
usage: python scripts/host_e2e.py [tuples_per_snapshot]"""
import os, re, shutil, subprocess, sys, tempfile, time
sys.path.insert(0, os.getcwd())
import bench

tuples = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
t0 = time.time()
snaps, halos = bench.cpu_sample_pair(777, tuples)
work = tempfile.mkdtemp(prefix="metro_e2e_")
cfg = bench.prepare_ref_run(work, snaps)
size = sum(os.path.getsize(os.path.join(work, "input", f)) for f in os.listdir(os.path.join(work, "input")))
print(f"pair: {halos} halos, {sum(s.n_tuples for s in snaps)} bound IDs, {size/1e6:.0f} MB of AHF text (written in {time.time()-t0:.0f} s)", flush=True)

def run(cmd, env=None):
    t = time.time()
    r = subprocess.run(cmd, cwd=work, capture_output=True, text=True, env=env)
    return time.time() - t, r

out_ref = os.path.join(work, "output")
ref = bench.ref_binary()
t_ref, r = run([ref, cfg])
assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
fw, bw = bench.parse_ref_times(r.stdout)
ref_files = {f: open(os.path.join(out_ref, f)).read() for f in os.listdir(out_ref) if f.endswith(".mtree")}
print(f"reference (1 core): {t_ref:.1f} s wall, FindProgenitors fwd+bwd {fw+bw:.1f} s", flush=True)

out_ours = os.path.join(work, "output_b200")
os.makedirs(out_ours)
cfg2 = os.path.join(work, "config_b200.cfg")
text = open(cfg).read().replace(f"pathOutput = {out_ref}/", f"pathOutput = {out_ours}/")
open(cfg2, "w").write(text + "treeFlavour = zoom\nnPTypes = 2\nnoPType = 1\n")
ours = os.path.join(os.getcwd(), "metrocpp_b200", "MetroCPP-b200")
t_ours, r2 = run([ours, cfg2], env=dict(os.environ, METRO_DEBUG="1"))
assert r2.returncode == 0, r2.stdout[-2000:] + r2.stderr[-2000:]
print(f"MetroCPP-b200: {t_ours:.2f} s wall (incl. process start, CUDA context, library load)")
for ln in r2.stdout.splitlines():
    if re.search(r"Files read|progenitors|Memory cleared", ln):
        print("   ", ln)
for ln in r2.stderr.splitlines():
    if "ingest" in ln:
        print("   ", ln)
ours_files = {f: open(os.path.join(out_ours, f)).read() for f in os.listdir(out_ours) if f.endswith(".mtree")}
assert ref_files.keys() == ours_files.keys() and ref_files, (ref_files.keys(), ours_files.keys())
for f in ref_files:
    assert ref_files[f] == ours_files[f], f
print(f".mtree files byte-identical: {sorted(ref_files)} ({sum(len(v) for v in ref_files.values())} bytes); wall-time ratio {t_ref / t_ours:.1f}x")
shutil.rmtree(work, ignore_errors=True)
