"""world_size-2 gloo test of the frame-sharded path: sharding covers every frame once, the
all-gather returns the same records a single process produces."""
import os
import subprocess
import sys

import numpy as np

from conftest import ROOT, PKG

WORKER = r"""
import os, sys
sys.path.insert(0, sys.argv[1])
import numpy as np, torch
from b200calib import dist as D
from b200calib.pipeline import FrameResult
rank, world, _ = D.init_from_env(backend="gloo")
n = 11
weights = [(7 * f) % 5 + 1 for f in range(n)]
mine = D.shard_frames(n, rank, world, weights if os.environ.get("BALANCE") else None)
res = []
for f in mine:
    r = FrameResult(segmentation=np.full(10 + f, 2, dtype=np.uint8))
    if f % 3:
        r.ee_pose = np.arange(7) + f
        r.icp_stats = np.array([0.5, 0.01 * f, 7.0, 100.0])
    res.append(r)
rec = D.gather_records(D.pack_records(mine, res))
if rank == 0:
    np.save(sys.argv[2], rec)
"""


def _expected(n=11):
    rec = np.full((n, 20), np.nan)
    for f in range(n):
        rec[f, 0], rec[f, 1], rec[f, 2] = f, float(f % 3 != 0), 10 + f
        if f % 3:
            rec[f, 3:10] = np.arange(7) + f
            rec[f, 17:20] = [0.5, 0.01 * f, 7.0]
    return rec


def _run(world, balance, tmp_path):
    out = str(tmp_path / f"rec_{world}_{balance}.npy")
    worker = tmp_path / "worker.py"
    worker.write_text(WORKER)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT=str(29600 + world + 10 * balance))
    if balance:
        env["BALANCE"] = "1"
    procs = []
    for r in range(world):
        e = dict(env, RANK=str(r), WORLD_SIZE=str(world), LOCAL_RANK=str(r))
        procs.append(subprocess.Popen([sys.executable, str(worker), PKG, out], env=e, stdout=subprocess.PIPE,
                                      stderr=subprocess.STDOUT, text=True))
    for p in procs:
        o, _ = p.communicate(timeout=300)
        assert p.returncode == 0, o[-2000:]
    return np.load(out)


def test_gather_world2_equals_single(tmp_path):
    want = _expected()
    for balance in (0, 1):
        got = _run(2, balance, tmp_path)
        assert np.array_equal(np.isnan(got), np.isnan(want))
        assert np.allclose(np.nan_to_num(got), np.nan_to_num(want))
    one = _run(1, 0, tmp_path)
    assert np.allclose(np.nan_to_num(one), np.nan_to_num(want))


def test_shard_frames_partition():
    sys.path.insert(0, PKG)
    from b200calib.dist import shard_frames
    for world in (1, 2, 4, 8):
        for weights in (None, [(13 * f) % 7 + 1 for f in range(37)]):
            parts = [shard_frames(37, r, world, weights) for r in range(world)]
            assert sorted(sum(parts, [])) == list(range(37))
            if weights:
                loads = [sum(weights[f] for f in p) for p in parts]
                assert max(loads) - min(loads) <= max(weights)
