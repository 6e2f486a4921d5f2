"""Device-timed match_pair of the 1024^3 synthetic pair with the per-stage CUDA-event breakdown (no e2e, no CPU leg)."""
import os, sys
sys.path.insert(0, os.getcwd())
import metrocpp_b200 as mb

halos = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
tuples = int(sys.argv[2]) if len(sys.argv) > 2 else 500_000_000
pid_bits = int(sys.argv[3]) if len(sys.argv) > 3 else 30
steps = int(sys.argv[4]) if len(sys.argv) > 4 else 3
prm = mb.make_params(nptypes=2, min_part_cmp=10, min_part_halo=100, fac_orphan_steps=100, max_orphan_steps=15, flavour=mb.FLAVOUR_GATHER)
spec = mb.SynthSpec(seed=20261019, n_halos=halos, n_tuples=tuples, pid_bits=pid_bits, shard=0, n_shards=1)
eng = mb.MetroEngine(0)
eng.synth_pair_device(prm, spec)
for i in range(steps):
    info = eng.match_pair(prm)
    t = eng.timings()
    print(f"step {i}: total {t['total']:.3f} ms path {t['path']} | A {t['pack_hist']:.3f} B {t['sort']:.3f} join {t['join']:.3f} "
          f"(C {t['sort_pass_ms'][2]:.3f} D {t['sort_pass_ms'][3]:.3f}) select {t['select']:.3f} | hits {info['n_hits']} pairs {info['n_pairs']} "
          f"N {sum(info['n_tuples'])} trees {info['n_trees']} tokens {info['n_tokens']} spilled {info['n_table_hits']} launches {t['total_launches']}", flush=True)
eng.close()
