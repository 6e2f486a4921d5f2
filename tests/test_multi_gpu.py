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
