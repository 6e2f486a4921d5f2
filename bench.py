#!/usr/bin/env python
"""bench.py -- progenitor matching of one snapshot pair (BASELINE.json metric).

    python bench.py --gpus 1 --steps 5 --warmup 3                 # our arm, 1 B200
    torchrun ... bench.py --gpus N ...                            # one rank per GPU (pid-range shards)
    python bench.py --impl reference ...                          # the reference's CPU path on the host cores

A "step" is ONE pass of the hot path over one synthetic snapshot pair: pack + two hash-partition
passes on the particle ID -> per-bucket shared-memory join (hit records out, partitioned by halo-A
range) -> shared-memory pair count -> selection (forward + backward trees, mutual-best filter,
orphan/token rule), device-resident tuples in, device-resident assignments out.  (Small pairs and
adversarial inputs take the LSD radix sort + run-length join path instead; same results.)
metric = bound particle IDs matched per second per pair = (N_A + N_B) / t_pair.

Workload at N=1: configs[1] of BASELINE.json, the synthetic 1024^3 full-box pair (1e6 halos,
5e8 bound IDs per snapshot, dense pids < 2^30).  N>1: the same per-GPU volume, sharded by
particle-ID range (weak scaling; N=8 is configs[2], the 2048^3 pair).
"""
from __future__ import annotations

import argparse
import json
import os
import re
import shutil
import statistics
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "bound particle IDs matched/sec per snapshot pair"
UNIT = "IDs/s"
FALLBACK_HBM_GBS = 6650.0          # B200_PROFILING.md fallback


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--halos", type=int, default=1_000_000, help="halos per snapshot per GPU")
    ap.add_argument("--tuples", type=int, default=500_000_000, help="bound IDs per snapshot per GPU")
    ap.add_argument("--pid-bits", type=int, default=0, help="0 = smallest even width that fits")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-sample", type=int, default=4_000_000, help="tuples per snapshot of the CPU sample (about 20 s of reference CPU time incl. its file ingest)")
    ap.add_argument("--workload", default="pair", choices=["pair", "chain20", "lgf54"],
                    help="pair: the headline 1024^3 pair (default).  chain20 / lgf54: BASELINE.json configs[4] / [3] - whole snapshot chains of distinct "
                         "synthetic catalogues through upload -> match -> fetch -> shift, one GPU")
    ap.add_argument("--chain-scale", type=float, default=1.0, help="scale factor on the halo and tuple counts of a chain workload")
    ap.add_argument("--sweep", default="", help="chain workloads, tuning only: NAME=v1,v2,... runs the chain once per value of that library "
                                                "environment switch (stderr: one line per value) before the measured default run")
    ap.add_argument("--ref-sample", type=int, default=400_000)
    ap.add_argument("--ref-ids", type=int, default=10_000_000, help="bound IDs per snapshot of the pair the multi-rank reference arm matches per step")
    return ap.parse_args()


# --------------------------------------------------------------------------------------------
# clocks (B200_PROFILING.md recipe) sampled DURING the timed region
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20",
                                          "-i", str(self.idx)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self) -> dict:
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx = float(f[2])
            except ValueError:
                continue
            for name, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


def measured_hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)"


def ncu_traffic():
    """DRAM bytes per unit (key or hit) of each kernel from the committed ncu capture (profiles/traffic.json)."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            return json.load(f)
    except Exception:
        return {}


def pick_pid_bits(tuples_total_one_snapshot: int) -> int:
    # ranks ~ tuples/1.15; the generator needs 2 x ranks <= 2^bits: the smallest ID range in which about half of the particles are
    # bound (BASELINE.json configs: dense IDs 1..2^30 for 5e8 bound IDs, 1..2^33 for 4e9)
    need = 2 * int(tuples_total_one_snapshot / 1.10) + 16
    return max(8, need.bit_length())


# --------------------------------------------------------------------------------------------
# CPU arm helpers: reference binary (oracle/_ref, built from /root/reference in the build
# container; the binaries travel with the repo snapshot) or the C restatement.
def host_catalogue_to_snapshot(cat, number, redshift):
    """lib host catalogue -> AHF snapshot (tuples grouped by halo, empty halos dropped: the
    reference's _particles parser cannot represent a 0-particle halo)."""
    import numpy as np
    from oracle.ahf_io import Snapshot
    order = np.argsort(cat.halo, kind="stable")
    halo = cat.halo[order]
    counts = np.bincount(halo, minlength=cat.n_halos)
    keep = counts > 0
    offsets = np.zeros(int(keep.sum()) + 1, np.int64)
    np.cumsum(counts[keep], out=offsets[1:])
    typ = cat.type[order] if cat.type is not None else np.ones(halo.shape[0], np.uint8)
    return Snapshot(number, redshift, cat.halo_id[keep], cat.npart.sum(axis=1)[keep].astype(np.int32), offsets,
                    cat.pid[order], typ)


def cpu_sample_pair(seed, tuples):
    import metrocpp_b200 as mb
    halos = max(50, tuples // 500)
    prm = mb.make_params(nptypes=2)
    spec = mb.SynthSpec(seed=seed, n_halos=halos, n_tuples=tuples, pid_bits=pick_pid_bits(tuples), shard=0, n_shards=1)
    c0, c1 = mb.synth_pair_host(prm, spec, 0), mb.synth_pair_host(prm, spec, 1)
    return [host_catalogue_to_snapshot(c1, 1, 0.100), host_catalogue_to_snapshot(c0, 2, 0.000)], halos


def ref_binary():
    p = os.path.join(ROOT, "oracle", "_ref", "MetroCPP_zoom")
    return p if os.path.isfile(p) and os.access(p, os.X_OK) else None


def prepare_ref_run(work, snaps, chunks=0, ngrid=24):
    """Lay out one reference run.  The snapshot lists the reference normally obtains from
    scripts/find_*.sh are pre-seeded in tmp/output_*.tmp (it reuses them when present,
    reference src/IOSettings.cpp:253-262), so nothing under /root/reference is needed."""
    from oracle import ahf_io, run_ref
    metro, inp, out = (os.path.join(work, d) for d in ("metro", "input", "output"))
    for d in (os.path.join(metro, "tmp"), os.path.join(metro, "data"), os.path.join(metro, "scripts"), inp, out):
        os.makedirs(d, exist_ok=True)
    ss = sorted(snaps, key=lambda s: s.number)
    for s in ss:
        if chunks:
            # full-box layout: one file pair per spatial sub-volume (compact block); task r reads chunks [r * nChunks/totTask, ...)
            from oracle import synth_np
            for c, part in enumerate(synth_np.split_slabs(s, chunks)):
                ahf_io.write_snapshot(inp, part, chunk=c, with_types=False)
        else:
            ahf_io.write_snapshot(inp, s, with_types=False)
    with open(os.path.join(metro, "tmp", "output_n.tmp"), "w") as f:
        f.write(f"{len(ss)}\n")
    with open(os.path.join(metro, "tmp", "output_id.tmp"), "w") as f:
        f.write("".join(f"{s.number:03d}\n" for s in ss))
    with open(os.path.join(metro, "tmp", "output_z.tmp"), "w") as f:
        f.write("".join(f"{s.redshift:.3f}\n" for s in ss))
    # a(t) / P(k) tables: loaded into splines at start-up (the spline asserts on < 3 points) and never
    # evaluated by the matching path: placeholder tables suffice.
    with open(os.path.join(metro, "data", "a2t_5Myr_planck.dat"), "w") as f:
        f.write("".join(f"{0.01 + 0.099 * i:.6f}\n" for i in range(11)))
    with open(os.path.join(metro, "data", "pk_planck.dat"), "w") as f:
        f.write("".join(f"{1e-4 * 2 ** i:.8g} {100.0 + i:.6g}\n" for i in range(11)))
    cfg = os.path.join(work, "config.cfg")
    if chunks:
        run_ref.write_config(cfg, metro=metro, inp=inp, out=out, n_snaps=len(ss), fullbox=True,
                             params=dict(minPartHalo=100, minPartCmp=10, facOrphanSteps=100, maxOrphanSteps=15))      # config/fullbox_*.cfg
        with open(cfg) as f:
            text = f.read().replace("nChunks = 1", f"nChunks = {chunks}").replace("nGrid = 4", f"nGrid = {ngrid}")       # config/fullbox_17_11.cfg:41,66: nGrid = 24
        with open(cfg, "w") as f:
            f.write(text)
    else:
        run_ref.write_config(cfg, metro=metro, inp=inp, out=out, n_snaps=len(ss), fullbox=False,
                             params=dict(minPartHalo=30, minPartCmp=10, facOrphanSteps=15, maxOrphanSteps=None))
    return cfg


def parse_ref_times(stdout):
    fw = re.search(r"forwards\.\.\.done in ([0-9.eE+-]+)s", stdout)
    bw = re.search(r"backwards\.\.\.done in ([0-9.eE+-]+)s", stdout)
    if not fw or not bw:
        raise RuntimeError("could not parse the reference's timing lines:\n" + stdout[-2000:])
    return float(fw.group(1)), float(bw.group(1))


def run_ref_instances(cfgs, works):
    """Run len(cfgs) single-rank reference processes concurrently; returns [(fwd_s, bwd_s)]."""
    procs = [subprocess.Popen([ref_binary(), c], cwd=w, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
             for c, w in zip(cfgs, works)]
    outs = [p.communicate()[0] for p in procs]
    return [parse_ref_times(o) for o in outs]


def time_oracle_port(snaps):
    from oracle import binding
    prm = binding.Params(2, 10, 30, 15, 0, binding.FLAVOUR_ZOOM_DUP)
    ss = sorted(snaps, key=lambda s: -s.number)
    c0 = binding.catalogue_from_snapshot(ss[0], 2, True)
    c1 = binding.catalogue_from_snapshot(ss[1], 2, True)
    t0 = time.perf_counter()
    binding.match_pair(prm, c0, c1)
    return time.perf_counter() - t0


def cpu_baseline(sample_tuples):
    """One core, one bounded sample of the same synthetic workload."""
    snaps, halos = cpu_sample_pair(4242, sample_tuples)
    n_ids = sum(s.n_tuples for s in snaps)
    desc = f"synthetic pair, {halos} halos, {n_ids} bound IDs (same generator as the GPU workload)"
    if ref_binary():
        work = tempfile.mkdtemp(prefix="metro_cpu_")
        try:
            cfg = prepare_ref_run(work, snaps)
            (fw, bw), = run_ref_instances([cfg], [work])
        finally:
            shutil.rmtree(work, ignore_errors=True)
        return {"value": n_ids / (fw + bw), "unit": UNIT, "cores": 1, "kind": "reference",
                "sample": desc + "; compiled reference (ZOOM build, 1 rank), FindProgenitors fwd+bwd CPU seconds",
                "seconds": fw + bw}
    t = time_oracle_port(snaps)
    return {"value": n_ids / t, "unit": UNIT, "cores": 1, "kind": "port",
            "sample": desc + "; C restatement (oracle/metro_oracle.c), match_pair wall seconds", "seconds": t}


# --------------------------------------------------------------------------------------------
def mp_ref_binary():
    p = os.path.join(ROOT, "oracle", "_ref", "MetroCPP_gather_mp")
    return p if os.path.isfile(p) and os.access(p, os.X_OK) else None


def parse_mp_times(stdout):
    """Rank 0's phase timers of one pair (reference src/main.cpp:178-283): buffer exchange, FindProgenitors forwards and
    backwards (each ends in an MPI_Barrier, so rank 0's time is the slowest rank's), gather of the trees to rank 0."""
    pats = {"exchange": r"Buffer exchanged in ([0-9.eE+-]+)s", "forward": r"forwards\.\.\.done in ([0-9.eE+-]+)s",
            "backward": r"backwards\.\.\.done in ([0-9.eE+-]+)s", "gather": r"Merger Trees gathered in ([0-9.eE+-]+)s"}
    out = {}
    for k, pat in pats.items():
        m = re.search(pat, stdout)
        if not m:
            raise RuntimeError(f"could not parse the reference's '{k}' timer:\n" + stdout[-2000:])
        out[k] = float(m.group(1))
    return out


def run_reference_arm(args, rank):
    """The reference's own CPU implementation on the host cores.  Preferred form: its REAL multi-rank full-box path - the unmodified
    sources built with -DGATHER_TREES against the multi-process MPI shim (oracle/mpi_shim_mp), one rank per core, chunked AHF input
    written from a numpy generator (oracle/synth_np.py; the product library is not involved).  Fallback when that binary is absent:
    one single-rank ZOOM process per core, or the C restatement."""
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    cores = max(1, min(cores, 32))
    if mp_ref_binary():
        from oracle import mpirun_shim, synth_np
        ids_per_snap = args.ref_ids
        snaps = synth_np.synth_pair(1000, max(50, ids_per_snap // 500), ids_per_snap)
        ids = sum(s.n_tuples for s in snaps)
        work = tempfile.mkdtemp(prefix="metro_refmp_")
        rates, phases = [], []
        try:
            cfg = prepare_ref_run(work, snaps, chunks=cores)
            for step in range(args.warmup + args.steps):
                for f in os.listdir(os.path.join(work, "output")):
                    os.remove(os.path.join(work, "output", f))
                t0 = time.perf_counter()
                res = mpirun_shim.run(cores, [mp_ref_binary(), cfg], cwd=work, timeout=1800)
                if any(r.returncode != 0 for r in res):
                    raise RuntimeError("the multi-rank reference failed:\n" + res[0].stdout[-2000:])
                ph = parse_mp_times(res[0].stdout)
                if step >= args.warmup:
                    rates.append((ids / sum(ph.values()), time.perf_counter() - t0))
                    phases.append(ph)
        finally:
            shutil.rmtree(work, ignore_errors=True)
        value = statistics.mean(r for r, _ in rates)
        ms = statistics.mean(t for _, t in rates) * 1e3
        mean_ph = {k: statistics.mean(p[k] for p in phases) for k in phases[0]}
        sample = (f"the reference's multi-rank full-box build (GATHER_TREES, unmodified sources, multi-process MPI shim), {cores} ranks = {cores} host cores, "
                  f"{cores} chunk files per snapshot (compact sub-volumes, nGrid = 24 as config/fullbox_17_11.cfg), synthetic full-box pair of {ids_per_snap} bound IDs per snapshot from a numpy generator; "
                  f"rate = (N_A+N_B) / (buffer exchange + FindProgenitors forwards + backwards + tree gather) on rank 0's timers "
                  f"(its barriers make them the slowest rank's); file reading and .mtree writing are outside, as in the GPU arm")
        line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "u64", "data": "synthetic",
                "config": {"workload": "synthetic 1024^3 full-box snapshot pair (bounded CPU sample of it)", "sample_ids_per_step": ids,
                           "reference_phase_seconds": mean_ph, "ranks": cores},
                "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "reference", "sample": sample},
                "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line), flush=True)
        return
    works, cfgs, ids = [], [], 0
    kind = "reference" if ref_binary() else "port"
    sample_sets = []
    for i in range(cores):
        snaps, halos = cpu_sample_pair(1000 + i, args.ref_sample)
        sample_sets.append(snaps)
        ids += sum(s.n_tuples for s in snaps)
        if kind == "reference":
            w = tempfile.mkdtemp(prefix=f"metro_ref{i}_")
            works.append(w)
            cfgs.append(prepare_ref_run(w, snaps))
    rates = []
    try:
        for step in range(args.warmup + args.steps):
            t0 = time.perf_counter()
            if kind == "reference":
                times = run_ref_instances(cfgs, works)
                rate = sum(sum(s.n_tuples for s in ss) / (fw + bw) for ss, (fw, bw) in zip(sample_sets, times))
            else:
                import concurrent.futures as cf
                with cf.ThreadPoolExecutor(cores) as ex:
                    ts = list(ex.map(time_oracle_port, sample_sets))
                rate = sum(sum(s.n_tuples for s in ss) / t for ss, t in zip(sample_sets, ts))
            if step >= args.warmup:
                rates.append((rate, time.perf_counter() - t0))
    finally:
        for w in works:
            shutil.rmtree(w, ignore_errors=True)
    value = statistics.mean(r for r, _ in rates)
    ms = statistics.mean(t for _, t in rates) * 1e3
    sample = (f"{cores} concurrent single-rank instances (multi-rank binary not built), each a synthetic pair of "
              f"{args.ref_sample} bound IDs per snapshot; rate = sum over instances of (N_A+N_B)/(fwd+bwd seconds)")
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u64", "data": "synthetic",
            "config": {"workload": "synthetic 1024^3 full-box snapshot pair (bounded CPU sample of it)", "sample_ids_per_step": ids},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def bind_to_gpu_numa_node(local_rank):
    """Multi-GPU runs: run this rank (and first-touch its pinned host buffers) on the NUMA node its GPU hangs off, so that the
    8 ranks of a node do not all pull their host-to-device copies through one memory controller.  Best effort: returns the node or None."""
    try:
        import torch
        p = torch.cuda.get_device_properties(local_rank)
        bdf = f"{p.pci_domain_id:04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0"
        with open(f"/sys/bus/pci/devices/{bdf}/numa_node") as f:
            node = int(f.read().strip())
        if node < 0:
            return None
        with open(f"/sys/devices/system/node/node{node}/cpulist") as f:
            spec = f.read().strip()
        cpus = set()
        for part in spec.split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        allowed = os.sched_getaffinity(0) & cpus
        if allowed:
            os.sched_setaffinity(0, allowed)
            return node
    except Exception:
        pass
    return None


# --------------------------------------------------------------------------------------------
def sharded_parity_check(eng, rank, world, dist, dev):
    """N>1 only, before anything is timed: a small pair (1500 halos, 2.5e5 tuples per snapshot) sharded by particle-ID range over
    the same ranks, communicator and front end as the benchmark, its concatenated per-rank trees, tokens and shifted catalogue
    compared bit for bit with the C oracle on the unsharded data (the checker; never timed).  Raises on any mismatch."""
    import numpy as np
    import metrocpp_b200 as mb
    prm = mb.make_params(nptypes=2, max_orphan_steps=5, flavour=mb.FLAVOUR_GATHER)
    bits = 20
    spec = mb.SynthSpec(seed=31, n_halos=1500, n_tuples=250_000, pid_bits=bits, shard=0, n_shards=1)
    c0, c1 = mb.synth_pair_host(prm, spec, 0), mb.synth_pair_host(prm, spec, 1)
    c0.type = None; c1.type = None
    per = (1 << bits) // world
    lo, hi = per * rank, (1 << bits) + 1 if rank == world - 1 else per * (rank + 1)

    def shard(c):
        m = (c.pid >= lo) & (c.pid < hi)
        return mb.HostCatalogue(c.halo_id, c.npart, c.n_orphan_steps, c.is_token, c.pid[m], c.halo[m], None)
    eng.upload(0, prm, shard(c0))
    eng.upload(1, prm, shard(c1))
    res = eng.find_progenitors_and_clean(prm)
    path = eng.timings()["path"]
    eng.shift_halos_parts(prm)
    d0 = eng.fetch_catalogue(0)
    mine = {"tree_main": res.tree_main, "tree_orphan": res.tree_orphan, "tree_off": res.tree_off, "clean_prog": res.clean_prog,
            "clean_ncommon": res.clean_ncommon, "token": res.token_halo, "n_hits": res.info["n_hits"], "raw_prog0": res.raw_prog[0],
            "raw_merit0": res.raw_merit[0], "shift": (d0.halo_id, d0.is_token, d0.pid, d0.halo), "path": path}
    parts = [None] * world
    dist.all_gather_object(parts, mine)
    out = {"world": world, "ok": True, "front_end": "partition join" if path == 1 else "LSD sort", "halos": 1500,
           "tuples": int(c0.n_tuples + c1.n_tuples)}
    if rank == 0:
        from oracle import binding
        from tests.helpers import to_oracle_catalogue
        oprm = binding.Params(2, prm.min_part_cmp, prm.min_part_halo, prm.fac_orphan_steps, prm.max_orphan_steps, 0)
        o0, o1 = to_oracle_catalogue(c0), to_oracle_catalogue(c1)
        ores = binding.match_pair(oprm, o0, o1)
        cat = lambda k: np.concatenate([p[k] for p in parts])
        counts = np.concatenate([np.diff(p["tree_off"]) for p in parts])
        checks = [(cat("tree_main"), ores.tree_main), (cat("tree_orphan"), ores.tree_orphan), (np.concatenate([[0], np.cumsum(counts)]), ores.tree_off),
                  (cat("clean_prog"), ores.clean_prog), (cat("clean_ncommon"), ores.clean_ncommon), (cat("raw_prog0"), ores.raw_prog[0]),
                  (cat("raw_merit0"), ores.raw_merit[0])]
        for p in parts:
            checks.append((p["token"], ores.token_halo))
            checks.append((np.array([p["n_hits"]]), np.array([ores.n_hits])))
        exp = binding.shift(oprm, o0, o1, ores.token_halo)
        for p in parts:
            checks.append((p["shift"][0], exp.halo_id)); checks.append((p["shift"][1], exp.is_token))
        H = np.uint64(exp.n_halos)
        got = np.sort(np.concatenate([p["shift"][2] * H + p["shift"][3].astype(np.uint64) for p in parts]))
        checks.append((got, np.sort(exp.pid * H + exp.halo.astype(np.uint64))))
        out["ok"] = bool(all(a.shape == b.shape and np.array_equal(a, b) for a, b in checks))
        out["trees"] = int(ores.tree_main.shape[0]); out["tokens"] = int(ores.token_halo.shape[0]); out["checked_arrays"] = len(checks)
    flag = [out["ok"]]
    dist.broadcast_object_list(flag, src=0, device=dev)
    if not flag[0]:
        raise RuntimeError(f"sharded parity check against the oracle FAILED at world={world}")
    return out


def run_ours(args, rank, world, local_rank):
    import numpy as np
    import torch
    import metrocpp_b200 as mb

    dist = None
    if world > 1:
        import torch.distributed as dist_mod
        dist = dist_mod
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    numa_node = bind_to_gpu_numa_node(local_rank) if world > 1 else None

    H_glob, N_glob = args.halos * world, args.tuples * world
    pid_bits = args.pid_bits or pick_pid_bits(N_glob)
    prm = mb.make_params(nptypes=2, min_part_cmp=10, min_part_halo=100, fac_orphan_steps=100, max_orphan_steps=15,
                         flavour=mb.FLAVOUR_GATHER)          # config/fullbox_*.cfg settings
    spec = mb.SynthSpec(seed=20261019, n_halos=H_glob, n_tuples=N_glob, pid_bits=pid_bits, shard=rank, n_shards=world)
    if world > 1:
        os.environ.setdefault("METRO_PART_MIN_N", "20000")     # the parity pair is small: send it through the benchmarked front end too
    eng = mb.MetroEngine(local_rank)
    stream = torch.cuda.current_stream(dev)
    eng.set_stream(stream.cuda_stream)
    parity = None
    if world > 1:
        uid = [eng.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0, device=dev)
        eng.comm_init(rank, world, uid[0])
        parity = sharded_parity_check(eng, rank, world, dist, dev)
    eng.synth_pair_device(prm, spec)
    if world > 1:
        # catalogue npart of the earlier halos = sum over the pid shards
        c1 = eng.fetch_catalogue(1, tuples=False)
        t = torch.from_numpy(c1.npart.astype(np.int32)).to(dev)
        dist.all_reduce(t)
        eng.set_npart(1, t.cpu().numpy())
    n0 = eng.catalogue_sizes(0)[1]
    n1 = eng.catalogue_sizes(1)[1]
    ids_local = n0 + n1

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def step():
        return eng.match_pair(prm)

    for _ in range(args.warmup):
        info = step()
    # ---- timed region: K steps, device events on the launching stream
    sampler = ClockSampler(local_rank)
    barrier()
    sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    pass_ms, stage = [], {"pack_hist": 0.0, "sort": 0.0, "join": 0.0, "exchange": 0.0, "select": 0.0, "exchange_sendrecv": 0.0}
    launches = 0
    ev0.record(stream)
    for _ in range(args.steps):
        info = step()
        t = eng.timings()
        pass_ms += t["sort_pass_ms"][: t["sort_launches"]]
        for k in stage:
            stage[k] += t[k]
        launches += t["total_launches"]
    ev1.record(stream)
    barrier()
    clocks = sampler.stop()
    ms_total = ev0.elapsed_time(ev1)
    tt = torch.tensor([ms_total, float(ids_local)], dtype=torch.float64, device=dev)
    if dist is not None:
        mx = tt.clone(); dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = tt.clone(); dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        ms_total, ids_glob = float(mx[0]), float(sm[1])
    else:
        ids_glob = float(ids_local)
    ms_per_step = ms_total / args.steps
    value = ids_glob / (ms_per_step * 1e-3)

    # ---- roofline: every streaming kernel of the path against the measured HBM peak; the headline object
    # is the one that takes the most time.  Algorithmic bytes per launch:
    #   partition path  A: 12 B (+1 with types) read + 8 B written per tuple   B: 8 + 8 per key
    #                   C: 8 B per key read + 8 B per hit written               D: 8 B per hit read
    #   LSD path        onesweep pass: 8 + 8 per key
    peak, peak_src = measured_hbm_peak()
    traffic = ncu_traffic()
    path = int(t["path"])
    hits_local = float(info["n_hits"]) / world            # n_hits is the global count; shards are statistically even
    kernels = []
    if path == 1:
        per_step = [statistics.mean(pass_ms[i::4]) for i in range(4)]
        tuple_in = 12.0 + (1.0 if prm.nptypes > 2 else 0.0)
        spec_k = [("part_kernel<PACK> x2 (pass A: pack + level-1 partition, one launch per catalogue)", (tuple_in + 8.0) * ids_local, 2, "pass_a"),
                  ("part_kernel (pass B: level-2 partition)", 16.0 * ids_local, 1, "pass_b"),
                  ("bucket_join_direct_kernel (pass C: shared-memory join fed by TMA bulk copies, hit records out)", 8.0 * ids_local + 8.0 * hits_local, 1, "pass_c"),
                  ("hit_reduce_kernel (pass D: shared-memory pair count)", 8.0 * hits_local, 1, "pass_d")]
        for (name, nbytes, nl, key), ms in zip(spec_k, per_step):
            ach = nbytes / (ms * 1e-3) / 1e9
            kernels.append({"kernel": name, "ms_per_step": ms, "launches_per_step": nl, "algorithmic_bytes_per_step": nbytes,
                            "achieved": ach, "frac": ach / peak, "traffic": traffic.get(key + "_dram_bytes_per_step")})
    else:
        avg_pass_ms = statistics.mean(pass_ms) if pass_ms else float("nan")
        nbytes = 16.0 * ids_local
        ach = nbytes / (avg_pass_ms * 1e-3) / 1e9
        per_key = traffic.get("onesweep_dram_bytes_per_key")
        kernels.append({"kernel": "onesweep_kernel (one LSD pass over the packed u64 keys)", "ms_per_step": avg_pass_ms * info["sort_passes"],
                        "launches_per_step": info["sort_passes"], "algorithmic_bytes_per_step": nbytes * info["sort_passes"],
                        "achieved": ach, "frac": ach / peak, "traffic": per_key * ids_local * info["sort_passes"] if per_key else None})
    dom = max(kernels, key=lambda k: k["ms_per_step"])
    # SURVEY 8(d) contract figure: bytes a textbook LSD sort + sort-reduce would move for this pair
    P_sort = (info["pid_bits"] + 7) // 8
    Pp = (2 * info["halo_bits"] + 7) // 8
    contract_bytes = ids_local * (8 + 24 * P_sort) + ids_local * 12 + hits_local * 8 + hits_local * (16 + 16 * Pp)
    roofline = {"bound": "hbm", "kernel": dom["kernel"], "achieved": dom["achieved"], "peak": peak, "unit": "GB/s",
                "frac": dom["frac"], "peak_source": peak_src,
                "traffic": (dom["traffic"] / dom["launches_per_step"]) if dom["traffic"] else None,
                "algorithmic_bytes_per_launch": dom["algorithmic_bytes_per_step"] / dom["launches_per_step"],
                "avg_launch_ms": dom["ms_per_step"] / dom["launches_per_step"], "launches_per_step": dom["launches_per_step"],
                "path": "partition join" if path == 1 else "LSD radix sort + run-length join",
                "kernels": kernels,
                "stage_ms_per_step": {k: v / args.steps for k, v in stage.items()},
                "survey_contract": {"bytes_per_pair": contract_bytes, "achieved": contract_bytes / (ms_per_step * 1e-3) / 1e9,
                                    "frac": contract_bytes / (ms_per_step * 1e-3) / 1e9 / peak,
                                    "note": "SURVEY 8(d) algorithmic bytes of an LSD sort + sort-reduce pipeline / measured pair time; "
                                            "the partition join moves fewer real bytes, so this can exceed 1"}}

    # ---- e2e: the user-facing call sequence with HOST buffers: upload both catalogues from pinned
    # memory, match, fetch the clean trees.
    e2e = None
    if not args.no_e2e:
        e2e = measure_e2e(eng, prm, dev, dist, max(1, min(args.steps, 3)), ids_glob)

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64",
            "data": "synthetic",
            "config": {"workload": ("synthetic 1024^3 full-box snapshot pair" if world == 1 else
                                    f"synthetic full-box snapshot pair sharded by particle-ID range over {world} GPUs"),
                       "halos_per_snapshot": H_glob, "bound_ids_per_snapshot": int(ids_glob / 2), "pid_bits": info["pid_bits"],
                       "front_end": roofline["path"], "l2": "inputs (>= 8 GB of keys per pass) far exceed the 126 MB L2",
                       "params": "minPartCmp=10 minPartHalo=100 facOrphanSteps=100 maxOrphanSteps=15 (fullbox cfg)",
                       "n_hits": info["n_hits"], "n_pairs": info["n_pairs"], "n_trees": info["n_trees"],
                       "n_tokens": info["n_tokens"], "n_table_hits": info["n_table_hits"],
                       "exchange_bytes_rank0": info["exchange_bytes"],
                       "exchange": ("none (single GPU)" if world == 1 else
                                    "partial pair counts -> owner(haloA)/owner(haloB): one grouped ncclSend/ncclRecv all-to-all "
                                    "+ all-reduce(max) of bestDescendant and token flags")},
            "clocks": clocks, "gpu_launches": launches, "roofline": roofline}
    if world > 1:
        # NVLink: bytes this rank put on the wire in the grouped send/recv all-to-all over the device time of that group alone
        # (CUDA events around ncclGroupStart/End; it includes waiting for the slowest peer, so it is a lower bound of the link rate)
        xb, xms = float(info["exchange_bytes"]), stage["exchange_sendrecv"] / args.steps
        line["exchange"] = {"bytes_sent_rank0": xb, "ms_sendrecv": xms, "ms_stage": stage["exchange"] / args.steps,
                            "nvlink_gbs": (xb / (xms * 1e-3) / 1e9) if xms > 0 else None,
                            "frac_of_900": (xb / (xms * 1e-3) / 1e9 / 900.0) if xms > 0 else None,
                            "frac_of_measured_770": (xb / (xms * 1e-3) / 1e9 / 770.0) if xms > 0 else None,
                            "note": "partial pair counts (12 B each) to owner(haloA) and owner(haloB): latency- and skew-bound, not bandwidth-bound"}
    if e2e:
        line["e2e"] = e2e
    if parity:
        line["parity_check"] = parity
    if world > 1:
        line["config"]["numa_node_rank0"] = numa_node
    if rank == 0:
        if world == 1 and not args.no_cpu_baseline:
            try:
                line["cpu_baseline"] = cpu_baseline(args.cpu_sample)
            except Exception as ex:                      # never lose the GPU line over the CPU leg
                line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": 1, "kind": "port", "sample": f"failed: {ex}"}
        print(json.dumps(line), flush=True)
    eng.close()
    if dist is not None:
        dist.destroy_process_group()


def measure_e2e(eng, prm, dev, dist, steps, ids_glob):
    """The user-facing call sequence with HOST buffers, every step: both catalogues go up from pinned host
    memory in the layout an _particles file / locParts holds them (particle IDs grouped by halo + one offset
    per halo: metro_catalogue_upload_csr), the pair is matched, the clean trees come back."""
    import numpy as np
    import torch
    import metrocpp_b200 as mb
    stream = torch.cuda.current_stream(dev)
    cats = []
    keep_pinned = []
    # the results land in a pinned arena (re-used every step)
    arena = torch.empty(256 << 20, dtype=torch.uint8, pin_memory=True)
    arena_pos = [0]

    def pinned_alloc(shape, dtype):
        n = int(np.prod(shape)) * np.dtype(dtype).itemsize
        lo = (arena_pos[0] + 63) & ~63
        if lo + n > arena.numel():
            return np.zeros(shape, dtype)
        arena_pos[0] = lo + n
        return arena.numpy()[lo:lo + n].view(dtype).reshape(shape)
    for slot in range(2):
        H, N = eng.catalogue_sizes(slot)
        halos = eng.fetch_catalogue(slot, tuples=False)
        # the halo table in pinned memory too (metro_catalogue: "host pointers, pinned preferred")
        for name in ("halo_id", "npart", "n_orphan_steps", "is_token"):
            a = getattr(halos, name)
            t = torch.empty(a.shape, dtype={"uint64": torch.int64, "int32": torch.int32, "uint8": torch.uint8}[a.dtype.name], pin_memory=True)
            v = t.numpy().view(a.dtype)
            v[...] = a
            setattr(halos, name, v)
            keep_pinned.append(t)
        pid = torch.empty(N, dtype=torch.int64, pin_memory=True)
        halo = torch.empty(N, dtype=torch.int32, pin_memory=True)
        rc = eng.lib.metro_catalogue_fetch_tuples(eng.h, slot, pid.data_ptr(), halo.data_ptr(), None)
        eng._check(rc)
        # group the tuples by halo (file order of an AHF _particles file); preparation, not timed
        h_dev = halo.to(dev)
        h_sorted, order = torch.sort(h_dev, stable=True)
        pid.copy_(pid.to(dev)[order])
        counts = torch.bincount(h_sorted, minlength=H)
        off = torch.zeros(H + 1, dtype=torch.int64, pin_memory=True)
        off[1:].copy_(torch.cumsum(counts, 0))
        del h_dev, h_sorted, order, counts
        torch.cuda.empty_cache()
        # IDs as 32-bit offsets from the smallest one when the range allows it (metro_catalogue_upload_csr32): what a reader holds
        # for a box of up to 1600^3 particles or for one rank's particle-ID range; halves the bytes over the bus
        pn = pid.numpy().view(np.uint64)
        base, top = int(pn.min()), int(pn.max())
        p32 = None
        if top - base < (1 << 32):
            p32 = torch.empty(N, dtype=torch.int32, pin_memory=True)
            np.subtract(pn, np.uint64(base), out=p32.numpy().view(np.uint32), casting="unsafe")
        cats.append((halos, pid, off, p32, base))
    narrow = all(c[3] is not None for c in cats)

    def h2d_bytes(use32):
        return sum((p32.numel() * 4 if use32 else p.numel() * 8) + o.numel() * 8 + c.n_halos * (8 + 4 * prm.nptypes + 4 + 4 + 1) for c, p, o, p32, _ in cats)

    def one(use32=narrow):
        for slot, (c, pid, off, p32, base) in enumerate(cats):
            hc = mb.HostCatalogue(c.halo_id, c.npart, c.n_orphan_steps, c.is_token, None if use32 else pid.numpy().view(np.uint64), None, None)
            if use32:
                eng.upload(slot, prm, hc, halo_offset=off.numpy(), pid32=p32.numpy().view(np.uint32), pid_base=base)
            else:
                eng.upload(slot, prm, hc, halo_offset=off.numpy())
        info = eng.match_pair(prm)
        arena_pos[0] = 0
        res = eng.fetch_result(info, raw=False, alloc=pinned_alloc)
        d2h = res.tree_main.nbytes + res.tree_orphan.nbytes + res.tree_off.nbytes + res.clean_prog.nbytes + \
            res.clean_ncommon.nbytes + res.token_halo.nbytes
        return d2h, info

    def timed(n_steps, use32):
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize(dev)
        t0 = time.perf_counter()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record(stream)
        for _ in range(n_steps):
            d2h, info = one(use32)
        ev1.record(stream)
        torch.cuda.synchronize(dev)
        wall = time.perf_counter() - t0
        ms = max(ev0.elapsed_time(ev1), wall * 1e3)         # host work (packing structs) counts too
        if dist is not None:
            t = torch.tensor([ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t[0])
        return ms, d2h, info

    one()                                                   # warm-up (allocations)
    ms, d2h, info = timed(steps, narrow)
    tt = eng.timings()                                      # device timings of the last match of the timed steps
    match_dev = {"device_ms": tt["total"], "path": tt["path"], "catalogue0_partitioned_under_upload": bool(tt["resident"] & 1), "catalogue1_pass_a_under_upload": bool(tt["resident"] & 2),
                 "pass_abcd_ms": [round(x, 3) for x in tt["sort_pass_ms"][:4]], "select_ms": round(tt["select"], 3)}
    wide = None
    if narrow:                                              # the same with full 64-bit IDs over the bus, for reference
        one(False)
        ms64, _, _ = timed(1, False)
        wide = {"ms_per_step": ms64, "value": ids_glob / (ms64 * 1e-3), "h2d_bytes_per_step": int(h2d_bytes(False))}
        one(narrow)
    h2d = h2d_bytes(narrow)
    # ---- chain step (what the reference's main loop does per older snapshot, src/main.cpp:154-308): metro_shift
    # (catalogue 1 -> 0 on the device, token tuples appended), upload of ONE new catalogue, match, fetch.  The two
    # synthetic catalogues alternate as "the next earlier snapshot", so every step joins the same number of IDs.
    chain = None
    try:
        n_chain = 4
        torch.cuda.synchronize(dev)
        c_ms, c_dev, c_res, c_h2d = [], [], [], 0
        for k in range(n_chain + 1):                         # first one warms the allocations of the shift
            slot_src = k % 2                                 # catalogue that becomes the new slot 1 (0 = A, 1 = B)
            c, pid, off, p32, base = cats[slot_src]
            t0c = time.perf_counter()
            eng.shift_halos_parts(prm)
            hc = mb.HostCatalogue(c.halo_id, c.npart, c.n_orphan_steps, c.is_token, None if narrow else pid.numpy().view(np.uint64), None, None)
            if narrow:
                eng.upload(1, prm, hc, halo_offset=off.numpy(), pid32=p32.numpy().view(np.uint32), pid_base=base)
            else:
                eng.upload(1, prm, hc, halo_offset=off.numpy())
            info_c = eng.match_pair(prm)
            eng.fetch_result(info_c, raw=False)
            torch.cuda.synchronize(dev)
            if k > 0:
                c_ms.append((time.perf_counter() - t0c) * 1e3)
                tt = eng.timings()
                c_dev.append(tt["total"]); c_res.append((tt["path"], tt["resident"] & 1, [round(x, 2) for x in tt["sort_pass_ms"][:4]], round(tt["select"], 2)))
                c_h2d = (p32.numel() * 4 if narrow else pid.numel() * 8) + off.numel() * 8
        ids_step = sum(info_c["n_tuples"])
        chain = {"value": ids_step / (statistics.median(c_ms) * 1e-3), "unit": UNIT, "ms_per_step": statistics.median(c_ms),
                 "ms_all_steps": [round(x, 1) for x in c_ms], "device_match_ms": statistics.median(c_dev),
                 "note": "median over the steps; a step re-plans (resident = 0) when catalogue 0, grown by the token tuples this ping-pong chain accumulates, exceeds the 1/8 head room of the plan",
                 "steps_path_resident_passABCD_select_ms": c_res, "h2d_bytes_per_step": int(c_h2d), "steps": n_chain,
                 "ids_per_step": int(ids_step), "per_rank": True,
                 "path": "metro_shift -> metro_catalogue_upload_csr (one catalogue) -> metro_match_pair -> metro_result_fetch"}
    except Exception as ex:                                  # informational: never lose the headline over it
        chain = {"failed": str(ex)}
    return {"value": ids_glob / (ms / steps * 1e-3), "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
            "ms_per_step": ms / steps, "steps": steps, "n_trees": info["n_trees"], "chain_step": chain,
            "h2d_gbs": h2d / (ms / steps * 1e-3) / 1e9, "ids_over_the_bus": "u32 offsets from the smallest ID" if narrow else "u64",
            "u64_ids": wide, "match": match_dev,
            "path": ("metro_catalogue_upload_csr32 x2 (pinned host: particle IDs grouped by halo as u32 offsets + i64 halo offsets; chunked copy, "
                     "widened and range-checked on the device under the copy)" if narrow else
                     "metro_catalogue_upload_csr x2 (pinned host: u64 particle IDs grouped by halo + i64 offsets)") +
                    " -> metro_match_pair -> metro_result_fetch (clean trees)"}


CHAIN_WORKLOADS = {
    # BASELINE.json configs[4]: full-box 20-snapshot chain with token / orphan tracking (config/fullbox_*.cfg settings, GATHER_TREES
    # flavour); sizes from SURVEY.md section 8d
    "chain20": dict(n_snaps=20, halos=100_000, tuples=50_000_000, multi=False, nptypes=2, flavour="gather",
                    params=dict(min_part_cmp=10, min_part_halo=100, fac_orphan_steps=100, max_orphan_steps=15),
                    text="full-box 20-snapshot chain, ~1e5 halos / ~5e7 bound IDs per snapshot, tokens on (minPartHalo=100 facOrphanSteps=100 maxOrphanSteps=15)"),
    # BASELINE.json configs[3]: multi-species Local Group zoom, 54 snapshots (config/lgf_gas.cfg settings, ZOOM build with NPTYPES=6)
    "lgf54": dict(n_snaps=54, halos=20_000, tuples=10_000_000, multi=True, nptypes=6, flavour="zoom",
                  params=dict(min_part_cmp=10, min_part_halo=30, fac_orphan_steps=50, max_orphan_steps=0),
                  text="multi-species (gas / DM / stars, NPTYPES=6) zoom chain of 54 snapshots, ~2e4 halos / ~1e7 bound IDs per snapshot, 2 % of the gas turns "
                       "into stars per step (minPartHalo=30 facOrphanSteps=50, maxOrphanSteps unset)"),
}


def run_chain_workload(args):
    """A whole chain of DISTINCT snapshots through the calls the reference's main loop makes per older snapshot (src/main.cpp:154-308):
    upload of the next catalogue from pinned host memory (CSR, u32 ID offsets when they fit), match (resident buckets of the previous
    pair where the layout allows), fetch of the clean trees, shift.  Per-step device and end-to-end times; one JSON line."""
    import numpy as np
    import torch
    import metrocpp_b200 as mb
    from oracle import synth_np                              # (the numpy input generator; no checker involved in this leg)
    w = CHAIN_WORKLOADS[args.workload]
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    halos, tuples = max(50, int(w["halos"] * args.chain_scale)), max(5000, int(w["tuples"] * args.chain_scale))
    t_gen = time.perf_counter()
    snaps = synth_np.synth_chain(20261020, w["n_snaps"], halos, tuples, multi_species=w["multi"])
    t_gen = time.perf_counter() - t_gen
    prm = mb.make_params(nptypes=w["nptypes"], flavour=mb.FLAVOUR_GATHER if w["flavour"] == "gather" else mb.FLAVOUR_ZOOM_DUP, **w["params"])
    base = int(min(int(s.pid.min()) for s in snaps))
    narrow = max(int(s.pid.max()) for s in snaps) - base < (1 << 32)
    cats = []
    for s in sorted(snaps, key=lambda x: -x.number):         # latest first
        H = s.n_halos
        npart = np.zeros((H, w["nptypes"]), np.int32)
        npart[:, 1] = s.npart                                 # what IOSettings::ReadHalos fills (src/IOSettings.cpp:961-989)
        ids = torch.empty(s.n_tuples, dtype=torch.int32 if narrow else torch.int64, pin_memory=True)
        if narrow:
            np.subtract(s.pid, np.uint64(base), out=ids.numpy().view(np.uint32), casting="unsafe")
        else:
            ids.numpy().view(np.uint64)[:] = s.pid
        typ = None
        if w["multi"]:
            typ = torch.empty(s.n_tuples, dtype=torch.uint8, pin_memory=True)
            typ.numpy()[:] = s.ptype
        off = torch.empty(H + 1, dtype=torch.int64, pin_memory=True)
        off.numpy()[:] = s.offsets
        cats.append((mb.HostCatalogue(np.ascontiguousarray(s.halo_id, np.uint64), npart, np.zeros(H, np.int32), np.zeros(H, np.uint8), None, None,
                                      None if typ is None else typ.numpy()), ids, off, s.n_tuples))
    del snaps
    eng = mb.MetroEngine(0)
    stream = torch.cuda.current_stream(dev)
    eng.set_stream(stream.cuda_stream)

    def upload(slot, c):
        hc, ids, off, _ = c
        if narrow:
            eng.upload(slot, prm, hc, halo_offset=off.numpy(), pid32=ids.numpy().view(np.uint32), pid_base=base)
        else:
            hc2 = mb.HostCatalogue(hc.halo_id, hc.npart, hc.n_orphan_steps, hc.is_token, ids.numpy().view(np.uint64), None, hc.type)
            eng.upload(slot, prm, hc2, halo_offset=off.numpy())

    def run_chain():
        steps = []
        upload(0, cats[0])
        for k in range(1, len(cats)):
            torch.cuda.synchronize(dev)
            t0 = time.perf_counter()
            upload(1, cats[k])
            info = eng.match_pair(prm)
            res = eng.fetch_result(info, raw=False)
            tm = eng.timings()
            eng.shift_halos_parts(prm)
            torch.cuda.synchronize(dev)
            steps.append({"wall_ms": (time.perf_counter() - t0) * 1e3, "device_ms": tm["total"], "path": tm["path"], "resident": tm["resident"] & 1, "pass_a_under_upload": bool(tm["resident"] & 2),
                          "ids": int(sum(info["n_tuples"])), "trees": int(info["n_trees"]), "tokens": int(info["n_tokens"]), "launches": tm["total_launches"],
                          "passes_ms": [round(x, 3) for x in tm["sort_pass_ms"][:4]], "select_ms": round(tm["select"], 3),
                          "h2d_bytes": cats[k][1].numel() * cats[k][1].element_size() + cats[k][2].numel() * 8 + (cats[k][3] if w["multi"] else 0)})
        return steps
    if args.sweep:                                           # tuning: the same chain under several values of one library switch
        name, vals = args.sweep.split("=")
        for v in vals.split(","):
            os.environ[name] = v
            run_chain()
            st = run_chain()
            mean = lambda f: round(statistics.mean(f(x) for x in st), 3)
            print(f"[sweep] {name}={v}: device {mean(lambda x: x['device_ms'])} ms/pair, wall {mean(lambda x: x['wall_ms'])}, passes A-D "
                  f"{[mean(lambda x, i=i: x['passes_ms'][i]) for i in range(4)]}, select {mean(lambda x: x['select_ms'])}, "
                  f"resident {sum(x['resident'] for x in st)}/{len(st)}, partition path {sum(x['path'] == 1 for x in st)}", file=sys.stderr, flush=True)
        os.environ.pop(name, None)
    run_chain()                                              # warm-up: allocations, plans
    sampler = ClockSampler(0)
    sampler.start()
    steps = run_chain()
    clocks = sampler.stop()
    ids = sum(s["ids"] for s in steps)
    wall = sum(s["wall_ms"] for s in steps)
    devms = sum(s["device_ms"] for s in steps)
    line = {"metric": METRIC, "value": ids / (devms * 1e-3), "unit": UNIT, "n_gpus": 1, "steps": len(steps), "warmup": len(steps),
            "ms_per_step": devms / len(steps), "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
            "config": {"workload": f"{args.workload}: {w['text']}", "halos_latest": int(cats[0][0].n_halos), "bound_ids_latest": int(cats[0][3]),
                       "snapshots": len(cats), "scale": args.chain_scale, "generator_seconds": round(t_gen, 1),
                       "ids_over_the_bus": "u32 offsets from the smallest ID" if narrow else "u64",
                       "l2": "every catalogue is a distinct snapshot; keys per pass exceed the L2 at full scale"},
            "clocks": clocks, "gpu_launches": int(sum(s["launches"] for s in steps)),
            "e2e": {"value": ids / (wall * 1e-3), "unit": UNIT, "ms_per_step": wall / len(steps), "h2d_bytes_per_step": int(statistics.mean(s["h2d_bytes"] for s in steps)),
                    "d2h_bytes_per_step": None, "path": "per older snapshot: metro_catalogue_upload_csr[32] -> metro_match_pair -> metro_result_fetch -> metro_shift"},
            "chain": {"resident_steps": int(sum(s["resident"] for s in steps)), "partition_path_steps": int(sum(s["path"] == 1 for s in steps)),
                      "wall_ms_all": [round(s["wall_ms"], 2) for s in steps], "device_ms_all": [round(s["device_ms"], 3) for s in steps],
                      "tokens_all": [s["tokens"] for s in steps], "first_steps": steps[:3], "last_step": steps[-1]}}
    print(json.dumps(line), flush=True)
    eng.close()


def main():
    args = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference_arm(args, rank)
        return
    if args.workload != "pair":
        if world > 1:
            raise SystemExit("the chain workloads run on one GPU")
        run_chain_workload(args)
        return
    run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
