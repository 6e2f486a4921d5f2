"""Worker of the at-size multi-GPU parity test: launched by torchrun, one rank per GPU.

A synthetic pair with 1e6 halos and 1.2e8 tuples is generated ON THE DEVICES, one particle-ID shard per rank (the way bench.py does
it), matched through the sharded path (local partition join -> NCCL all-to-all of the partial pair counts -> owner-side selection);
the tuples are then collected on rank 0, and the concatenated per-rank trees and tokens are compared in full with the C oracle run
on the unsharded data."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import metrocpp_b200 as mb  # noqa: E402
from oracle import binding  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    dist.init_process_group("gloo")
    torch.cuda.set_device(local)
    halos, tuples, bits = 1_000_000, 50_000_000, 28
    prm = mb.make_params(nptypes=2, min_part_cmp=10, min_part_halo=100, fac_orphan_steps=100, max_orphan_steps=15, flavour=mb.FLAVOUR_GATHER)
    spec = mb.SynthSpec(seed=20261019, n_halos=halos, n_tuples=tuples, pid_bits=bits, shard=rank, n_shards=world)
    eng = mb.MetroEngine(local)
    uid = [eng.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    eng.comm_init(rank, world, uid[0])
    eng.synth_pair_device(prm, spec)
    # catalogue npart of the earlier halos = sum over the pid shards
    c1h = eng.fetch_catalogue(1, tuples=False)
    t = torch.from_numpy(c1h.npart.astype(np.int32))
    dist.all_reduce(t)
    eng.set_npart(1, t.numpy())
    res = eng.find_progenitors_and_clean(prm)
    tm = eng.timings()
    assert tm["path"] == 1, tm
    c0, c1 = eng.fetch_catalogue(0), eng.fetch_catalogue(1)
    mine = {"tree_main": res.tree_main, "tree_orphan": res.tree_orphan, "tree_off": res.tree_off, "clean_prog": res.clean_prog,
            "clean_ncommon": res.clean_ncommon, "token": res.token_halo, "info": res.info, "raw_off0": res.raw_off[0], "raw_prog0": res.raw_prog[0],
            "raw_nc0": res.raw_ncommon[0], "raw_merit0": res.raw_merit[0],
            "c0": (c0.halo_id, c0.npart, c0.pid, c0.halo), "c1": (c1.halo_id, c1.pid, c1.halo)}
    parts = [None] * world if rank == 0 else None
    dist.gather_object(mine, parts, dst=0)
    if rank == 0:
        def full(slot_key, npart):
            hid = parts[0][slot_key][0]
            pid = np.concatenate([p[slot_key][-2] for p in parts])
            halo = np.concatenate([p[slot_key][-1] for p in parts])
            H = hid.shape[0]
            return binding.Catalogue(np.ascontiguousarray(hid, np.uint64), np.ascontiguousarray(npart, np.int32), np.zeros(H, np.int32), np.zeros(H, np.uint8),
                                     np.ascontiguousarray(pid, np.uint64), np.ascontiguousarray(halo, np.uint32), np.ones(pid.shape[0], np.uint8))
        o0 = full("c0", parts[0]["c0"][1])
        o1 = full("c1", t.numpy())
        oprm = binding.Params(2, prm.min_part_cmp, prm.min_part_halo, prm.fac_orphan_steps, prm.max_orphan_steps, 0)
        ores = binding.match_pair(oprm, o0, o1)
        cat = lambda k: np.concatenate([p[k] for p in parts])
        counts = np.concatenate([np.diff(p["tree_off"]) for p in parts])
        np.testing.assert_array_equal(cat("tree_main"), ores.tree_main)
        np.testing.assert_array_equal(cat("tree_orphan"), ores.tree_orphan)
        np.testing.assert_array_equal(np.concatenate([[0], np.cumsum(counts)]), ores.tree_off)
        np.testing.assert_array_equal(cat("clean_prog"), ores.clean_prog)
        np.testing.assert_array_equal(cat("clean_ncommon"), ores.clean_ncommon)
        np.testing.assert_array_equal(sum(np.diff(p["raw_off0"]) for p in parts), np.diff(ores.raw_off[0]))
        np.testing.assert_array_equal(cat("raw_prog0"), ores.raw_prog[0])
        np.testing.assert_array_equal(cat("raw_nc0"), ores.raw_ncommon[0])
        np.testing.assert_array_equal(cat("raw_merit0"), ores.raw_merit[0])
        for p in parts:
            np.testing.assert_array_equal(p["token"], ores.token_halo)
            assert p["info"]["n_hits"] == ores.n_hits and p["info"]["n_untracked"] == ores.n_untracked
        print(f"MULTI_GPU_AT_SIZE_OK world={world} halos={o0.n_halos}/{o1.n_halos} tuples={o0.n_tuples}+{o1.n_tuples} hits={ores.n_hits} "
              f"trees={ores.tree_main.shape[0]} tokens={ores.token_halo.shape[0]} device_ms={tm['total']:.2f} exchange_bytes={[p['info']['exchange_bytes'] for p in parts]}", flush=True)
    eng.close()
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
