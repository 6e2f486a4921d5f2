"""Multi-GPU parity (needs >= 2 GPUs: run with `gpurun --gpus 2`): sharded match == oracle."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpu():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


@pytest.mark.gpu
@pytest.mark.parametrize("world,nptypes,front_end", [(2, 2, "lsd"), (2, 6, "lsd"), (2, 2, "partition"), (2, 6, "partition"), (4, 2, "partition")])
def test_sharded_match_equals_oracle(world, nptypes, front_end):
    if _ngpu() < world:
        pytest.skip(f"needs {world} GPUs")
    env = dict(os.environ)
    if front_end == "partition":
        env["METRO_PART_MIN_N"] = "20000"          # the shards are small: force the partition join front end
        env["METRO_EXPECT_PATH"] = "1"
    else:
        env["METRO_FORCE_LSD"] = "1"
        env["METRO_EXPECT_PATH"] = "0"
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(29700 + world + nptypes), os.path.join(ROOT, "tests", "multi_gpu_worker.py"), str(nptypes)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT, env=env)
    assert r.returncode == 0 and "MULTI_GPU_OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]


@pytest.mark.gpu
@pytest.mark.parametrize("world", [2, 4])
def test_host_program_ranks_concatenate_to_the_single_gpu_file(tmp_path, world):
    """`MetroCPP-b200 --ranks K cfg` (one process per GPU, every rank reads all files and keeps its particle-ID range, per-rank
    .mtree files like the reference's, src/IOSettings.cpp:1086): the rank files concatenated in rank order are byte-identical to the
    reference's single-task file, on the 20-snapshot token chain."""
    if _ngpu() < world:
        pytest.skip(f"needs {world} GPUs")
    from tests import golden_io
    from tests.test_host_program import BIN, write_case
    snaps, params, mtree = golden_io.load("fullbox_chain20")
    cfg, out = write_case(str(tmp_path), snaps, params, "gather", fullbox_names=True)
    r = subprocess.run([BIN, "--ranks", str(world), cfg], capture_output=True, text=True, timeout=900,
                       env=dict(os.environ, METRO_PART_MIN_N="1000"))
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    for number, text in mtree["gather"].items():
        got = "".join(open(os.path.join(out, f"test_{number:03d}.{rank}.mtree")).read() for rank in range(world))
        assert got == text, f"snapshot {number}"


@pytest.mark.gpu
@pytest.mark.parametrize("world", [2])
def test_sharded_match_at_size_equals_oracle(world):
    """1e6 halos / 1.2e8 tuples generated on the devices, one pid shard per GPU, through the sharded partition join; the concatenated
    per-rank raw and clean trees and the tokens compared in full with the oracle on the unsharded data (tests/multi_gpu_at_size_worker.py)."""
    if _ngpu() < world:
        pytest.skip(f"needs {world} GPUs")
    import torch
    if torch.cuda.get_device_properties(0).total_memory < 100e9:
        pytest.skip("needs 180 GB B200s")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(29760 + world), os.path.join(ROOT, "tests", "multi_gpu_at_size_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=1500, cwd=ROOT, env=dict(os.environ, METRO_DEBUG="1"))
    assert r.returncode == 0 and "MULTI_GPU_AT_SIZE_OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
    assert "partition join attempt overflowed" not in r.stderr, r.stderr[-2000:]
