"""b200calib/data_engine.py: mirrors of the reference's PickleDataEngine / PCDDataEngine (app/data_engine.py:53-204)
and the NumPy PCD reader that stands in for Open3D's. Host logic only."""
import json
import os
import pickle

import numpy as np
import pytest

from conftest import GOLDEN
from b200calib import data_engine as DE
from b200calib.ingest import pack_xyzrgb

REF_PCD = "/root/reference/app/hand_files/hand.pcd"


def _write_pcd(path, rec, mode):
    n = len(rec)
    head = ("# .PCD v0.7 - Point Cloud Data file format\nVERSION 0.7\nFIELDS x y z rgb\nSIZE 4 4 4 4\nTYPE F F F F\n"
            f"COUNT 1 1 1 1\nWIDTH {n}\nHEIGHT 1\nVIEWPOINT 0 0 0 1 0 0 0\nPOINTS {n}\nDATA {mode}\n")
    with open(path, "wb") as fp:
        fp.write(head.encode())
        if mode == "binary":
            fp.write(np.ascontiguousarray(rec, np.float32).tobytes())
        else:
            for r in rec:
                fp.write((" ".join(repr(float(v)) for v in r) + "\n").encode())


def _frame(rng, n):
    xyz = rng.normal(0, 1, (n, 3)).astype(np.float32)
    rgb = rng.integers(0, 256, (n, 3)).astype(np.uint8)
    xyz[rng.random(n) < 0.1] = np.nan
    return xyz, rgb


def test_pcd_reader_and_engine(tmp_path):
    rng = np.random.default_rng(3)
    frames = {}
    for name, mode in ((2, "binary"), (10, "ascii"), (1, "binary"), (33, "binary")):
        xyz, rgb = _frame(rng, 500 + name)
        rec = pack_xyzrgb(xyz, rgb)
        _write_pcd(tmp_path / f"{name}.pcd", rec, mode)
        pose = rng.normal(size=7).astype(np.float32)
        np.save(tmp_path / f"{name}_robot2ee_pose.npy", pose)
        np.save(tmp_path / f"{name}.npy", pose[::-1].copy())
        frames[name] = (xyz, rgb, rec, pose)
    for name, (xyz, rgb, rec, pose) in frames.items():
        got = DE.pcd_records(str(tmp_path / f"{name}.pcd"))
        assert np.array_equal(got.view(np.uint32), rec.view(np.uint32)), name       # bit-exact, NaNs included
        assert np.array_equal(DE.unpack_rgb(got[:, 3]), rgb.astype(np.float32) / np.float32(255.0))
    eng = DE.PCDDataEngine(str(tmp_path), cyclic=False, step=2)      # numeric order 1, 2, 10, 33 -> 1, 10
    assert [os.path.basename(p) for p in eng.data] == ["1.pcd", "10.pcd"] and len(eng) == 2
    for name in (1, 10):
        xyz, rgb, rec, pose = frames[name]
        f = eng.get()
        keep = np.isfinite(xyz).all(1)
        assert np.array_equal(f.points, xyz[keep]) and f.points.dtype == np.float32
        assert np.array_equal(f.rgb, rgb[keep].astype(np.float32) / np.float32(255.0))
        assert np.array_equal(f.ee2base_pose, np.concatenate((pose[:3], pose[6:7], pose[3:6])))   # xyzw -> wxyz
        assert f.gt_pose is None and f.id.endswith(f"{name}.pcd")
    assert eng.get() is None
    eng = DE.PCDDataEngine(str(tmp_path), cyclic=True, step=1)
    r, ee2base, path = eng.get_records()
    assert np.array_equal(r.view(np.uint32), frames[1][2].view(np.uint32)) and path.endswith("1.pcd")
    for _ in range(4):
        last = eng.get_records()
    assert last[2].endswith("1.pcd")                                   # cyclic


def test_pickle_engine(tmp_path):
    rng = np.random.default_rng(4)
    items = []
    content = {}
    for i, (name, pos, style) in enumerate(((12, "c2", "dict"), (3, "c2", "tuple"), (7, "c1", "dict"))):
        pts = rng.normal(size=(50, 3)).astype(np.float32)
        rgb = rng.random((50, 3)).astype(np.float32)
        pose = rng.normal(size=7).astype(np.float32)
        r2e = rng.normal(size=7).astype(np.float32)
        obj = (dict(points=pts, rgb=rgb, labels=np.zeros(50, np.float32), instance_labels=np.zeros(50, np.float32),
                    pose=pose, joint_angles=np.zeros(9, np.float32), robot2ee_pose=r2e)
               if style == "dict" else (pts, rgb, np.zeros(50, np.float32), np.zeros(50, np.float32), pose))
        os.makedirs(tmp_path / "d", exist_ok=True)
        with open(tmp_path / "d" / f"{name}.pickle", "wb") as fp:
            pickle.dump(obj, fp)
        items.append(dict(filepath=f"d/{name}.pickle", position=pos, light="c1", arm_point_count=1))
        content[name] = (pts, rgb, pose, r2e if style == "dict" else None)
    with open(tmp_path / "splits.json", "w") as fp:
        json.dump(dict(train=[], val=[], test=items), fp)
    eng = DE.PickleDataEngine(str(tmp_path / "splits.json"), split="test", cyclic=False)
    assert len(eng) == 3
    wxyz = lambda p: np.concatenate((p[:3], p[6:7], p[3:6]))  # noqa: E731
    for name in (7, 3, 12):                                           # (position, numeric name): c1/7, c2/3, c2/12
        pts, rgb, pose, r2e = content[name]
        f = eng.get()
        assert np.array_equal(f.points, pts) and np.array_equal(f.rgb, rgb)
        assert np.array_equal(f.gt_pose, wxyz(pose))
        assert (f.ee2base_pose is None) if r2e is None else np.array_equal(f.ee2base_pose, wxyz(r2e))
    assert eng.get() is None


@pytest.mark.skipif(not os.path.exists(REF_PCD), reason="reference checkout absent (GPU box)")
def test_reader_on_the_reference_cad_file():
    """the reference's own app/hand_files/hand.pcd (binary, x y z rgb) against the committed copy of its xyz."""
    rec = DE.pcd_records(REF_PCD)
    want = np.load(os.path.join(GOLDEN, "cad_hand_points.npz"))["xyz"]
    assert rec.shape == (4480, 4) and np.array_equal(rec[:, :3], want.astype(np.float32))
    rgb = DE.unpack_rgb(rec[:, 3])
    assert rgb.min() >= 0.0 and rgb.max() <= 1.0
