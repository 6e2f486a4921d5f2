"""Worker of the multi-GPU parity test: launched by torchrun, one rank per GPU.

Every rank uploads the FULL halo tables and ITS particle-ID range of the tuples, runs the sharded
match (local join -> NCCL all-to-all of partial pair counts -> owner-side selection), then rank 0
concatenates the per-rank clean trees and checks them, the tokens and the shifted catalogue against
the oracle run on the unsharded data.
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import metrocpp_b200 as mb  # noqa: E402
from oracle import binding  # noqa: E402
from tests.helpers import sorted_tuples, to_oracle_catalogue  # noqa: E402


def shard(cat, lo, hi):
    m = (cat.pid >= lo) & (cat.pid < hi)
    return mb.HostCatalogue(cat.halo_id, cat.npart, cat.n_orphan_steps, cat.is_token, cat.pid[m], cat.halo[m],
                            None if cat.type is None else cat.type[m])


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    dist.init_process_group("gloo")
    torch.cuda.set_device(local)
    nptypes = int(sys.argv[1]) if len(sys.argv) > 1 else 2
    prm = mb.make_params(nptypes=nptypes, max_orphan_steps=5, flavour=mb.FLAVOUR_GATHER)
    bits = 20
    spec = mb.SynthSpec(seed=31, n_halos=1500, n_tuples=250_000, pid_bits=bits, shard=0, n_shards=1)
    c0, c1 = mb.synth_pair_host(prm, spec, 0), mb.synth_pair_host(prm, spec, 1)
    if nptypes <= 2:
        c0.type = None; c1.type = None
    per = (1 << bits) // world
    lo, hi = per * rank, per * (rank + 1)
    eng = mb.MetroEngine(local)
    uid = [eng.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    eng.comm_init(rank, world, uid[0])
    eng.upload(0, prm, shard(c0, lo, hi))
    eng.upload(1, prm, shard(c1, lo, hi))
    res = eng.find_progenitors_and_clean(prm)
    if "METRO_EXPECT_PATH" in os.environ:
        assert eng.timings()["path"] == int(os.environ["METRO_EXPECT_PATH"]), eng.timings()
    parts = [None] * world
    dist.all_gather_object(parts, {"tree_main": res.tree_main, "tree_orphan": res.tree_orphan, "tree_off": res.tree_off,
                                   "clean_prog": res.clean_prog, "clean_ncommon": res.clean_ncommon, "token": res.token_halo,
                                   "info": res.info, "raw_off0": res.raw_off[0], "raw_prog0": res.raw_prog[0],
                                   "raw_nc0": res.raw_ncommon[0], "raw_merit0": res.raw_merit[0]})
    eng.shift_halos_parts(prm)
    d0 = eng.fetch_catalogue(0)
    shifted = [None] * world
    dist.all_gather_object(shifted, (d0.halo_id, d0.npart, d0.n_orphan_steps, d0.is_token, d0.pid, d0.halo, d0.type))
    ok = True
    if rank == 0:
        oprm = binding.Params(nptypes, prm.min_part_cmp, prm.min_part_halo, prm.fac_orphan_steps, prm.max_orphan_steps, 0)
        o0, o1 = to_oracle_catalogue(c0), to_oracle_catalogue(c1)
        ores = binding.match_pair(oprm, o0, o1)
        tree_main = np.concatenate([p["tree_main"] for p in parts])
        tree_orphan = np.concatenate([p["tree_orphan"] for p in parts])
        clean_prog = np.concatenate([p["clean_prog"] for p in parts])
        clean_nc = np.concatenate([p["clean_ncommon"] for p in parts])
        counts = np.concatenate([np.diff(p["tree_off"]) for p in parts])
        tree_off = np.concatenate([[0], np.cumsum(counts)])
        np.testing.assert_array_equal(tree_main, ores.tree_main)
        np.testing.assert_array_equal(tree_orphan, ores.tree_orphan)
        np.testing.assert_array_equal(tree_off, ores.tree_off)
        np.testing.assert_array_equal(clean_prog, ores.clean_prog)
        np.testing.assert_array_equal(clean_nc, ores.clean_ncommon)
        for p in parts:
            np.testing.assert_array_equal(p["token"], ores.token_halo)          # global on every rank
            assert p["info"]["n_hits"] == ores.n_hits and p["info"]["n_untracked"] == ores.n_untracked
        # forward raw trees: rank r holds the halos of its range, empty elsewhere
        raw_counts = sum(np.diff(p["raw_off0"]) for p in parts)
        np.testing.assert_array_equal(raw_counts, np.diff(ores.raw_off[0]))
        np.testing.assert_array_equal(np.concatenate([p["raw_prog0"] for p in parts]), ores.raw_prog[0])
        np.testing.assert_array_equal(np.concatenate([p["raw_merit0"] for p in parts]), ores.raw_merit[0])
        exp = binding.shift(oprm, o0, o1, ores.token_halo)
        for s in shifted:
            np.testing.assert_array_equal(s[0], exp.halo_id)
            np.testing.assert_array_equal(s[1], exp.npart)
            np.testing.assert_array_equal(s[2], exp.n_orphan_steps)
        pid = np.concatenate([s[4] for s in shifted]); halo = np.concatenate([s[5] for s in shifted])
        typ = np.concatenate([s[6] for s in shifted])
        assert sorted_tuples(pid, halo, typ) == sorted_tuples(exp.pid, exp.halo, exp.type)
        print(f"MULTI_GPU_OK world={world} trees={tree_main.shape[0]} tokens={ores.token_halo.shape[0]} "
              f"exchange_bytes={[p['info']['exchange_bytes'] for p in parts]}")
    eng.close()
    dist.barrier()
    dist.destroy_process_group()
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
