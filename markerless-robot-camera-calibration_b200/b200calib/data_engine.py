"""File readers that feed the batched pipeline: mirrors of the reference's data engines

  PickleDataEngine    app/data_engine.py:53-158   split JSON (dataset/sample_splits.json) + pickles in the README
                                                  schema (README.md:52-63), utils/file_utils.py:4-16
  PCDDataEngine       app/data_engine.py:161-204  a directory of <n>.pcd + <n>.npy (+ <n>_robot2ee_pose.npy)

Same constructor arguments, ordering and pose conventions (poses come back WXYZ: utils/transformation.py switch_w).
The reference reads .pcd files with Open3D (not installed here); `read_pcd` is a small NumPy reader for the PCD v0.7
files it uses (FIELDS x y z rgb, DATA ascii | binary; PCL packs rgb as 0x00RRGGBB in the bits of a float).
`PCDDataEngine.get_records()` returns the organised [N,4] (x, y, z, packed rgb) records `b200calib.ingest.ingest_clouds`
consumes, so a batch of frames goes file -> pinned host buffer -> device without a per-point NumPy pass; `get()` returns
the same frame the reference's `get()` returns (ROI / non-finite filter applied, colours in [0, 1]).
"""
import glob
import json
import os
import pickle
from dataclasses import dataclass
from datetime import datetime, timezone
from itertools import cycle
from typing import Optional

import numpy as np

from .transformation import switch_w

_NP_TYPES = {("F", 4): np.float32, ("F", 8): np.float64, ("U", 1): np.uint8, ("U", 2): np.uint16, ("U", 4): np.uint32,
             ("I", 1): np.int8, ("I", 2): np.int16, ("I", 4): np.int32}


def _utcnow():
    """naive UTC time stamp, as datetime.utcnow() gave the reference"""
    return datetime.now(timezone.utc).replace(tzinfo=None)


@dataclass
class PointCloud:
    """PointCloudDTO of app/dto.py:7-15."""
    points: np.ndarray
    rgb: np.ndarray
    timestamp: Optional[datetime] = None
    ee2base_pose: Optional[np.ndarray] = None
    joint_angles: Optional[np.ndarray] = None
    id: Optional[str] = None
    gt_pose: Optional[np.ndarray] = None


def read_pcd(path):
    """PCD v0.7 -> dict of per-field arrays (field name -> [N] or [N, count]). ascii and binary DATA."""
    with open(path, "rb") as fp:
        header = {}
        while True:
            line = fp.readline()
            if not line:
                raise ValueError(f"{path}: no DATA line")
            text = line.decode("ascii", "replace").strip()
            if not text or text.startswith("#"):
                continue
            key, _, val = text.partition(" ")
            header[key.upper()] = val.split()
            if key.upper() == "DATA":
                break
        fields = header["FIELDS"]
        sizes = [int(v) for v in header["SIZE"]]
        types = header["TYPE"]
        counts = [int(v) for v in header.get("COUNT", ["1"] * len(fields))]
        n = int(header["POINTS"][0]) if "POINTS" in header else int(header["WIDTH"][0]) * int(header["HEIGHT"][0])
        dtype = np.dtype([(f, _NP_TYPES[(t, s)], (c,)) if c > 1 else (f, _NP_TYPES[(t, s)])
                          for f, s, t, c in zip(fields, sizes, types, counts)])
        mode = header["DATA"][0].lower()
        if mode == "binary":
            rec = np.frombuffer(fp.read(n * dtype.itemsize), dtype=dtype, count=n)
        elif mode == "ascii":
            cols = np.loadtxt(fp, dtype=np.float64, ndmin=2)[:n]
            rec = np.zeros(n, dtype=dtype)
            c0 = 0
            for f, c in zip(fields, counts):
                block = cols[:, c0:c0 + c]
                if f == "rgb" and dtype[f] == np.float32:
                    # an ascii float field goes through float32 exactly as PCL writes it
                    rec[f] = block[:, 0].astype(np.float32)
                else:
                    rec[f] = block[:, 0] if c == 1 else block
                c0 += c
        else:
            raise NotImplementedError(f"{path}: DATA {mode} (binary_compressed is not used by the reference's files)")
    return {f: rec[f] for f in fields}


def pcd_records(path):
    """[N,4] float32 records (x, y, z, packed rgb) of a .pcd file, organised order, nothing filtered."""
    d = read_pcd(path)
    out = np.empty((len(d["x"]), 4), dtype=np.float32)
    out[:, 0], out[:, 1], out[:, 2] = d["x"], d["y"], d["z"]
    rgb = d.get("rgb", d.get("rgba"))
    out[:, 3] = rgb if rgb is not None and rgb.dtype == np.float32 else (
        np.zeros(len(out), np.float32) if rgb is None else rgb.astype(np.uint32).view(np.float32))
    return out


def unpack_rgb(packed):
    """PCL packed rgb (float32 bit pattern 0x00RRGGBB) -> [N,3] float32 in [0, 1] (what Open3D's reader returns)."""
    u = np.ascontiguousarray(packed, dtype=np.float32).view(np.uint32)
    return np.stack(((u >> 16) & 255, (u >> 8) & 255, u & 255), axis=1).astype(np.float32) / np.float32(255.0)


def _roi_mask(points, lim=500.0):
    """utils/data.py:58-75 get_roi_mask with its default limits: only non-finite / absurd points fall outside."""
    m = np.ones(len(points), dtype=bool)
    for d in range(3):
        m &= (points[:, d] > -lim) & (points[:, d] < lim)
    return m


class PickleDataEngine:
    """app/data_engine.py:53-158. data_path: split JSON; frames sorted by (position, numeric file name)."""

    def __init__(self, data_path, split="test", cyclic=True, base_path=None):
        self.data = {split: []}
        with open(data_path, "r") as fp:
            self.data.update(json.load(fp))
        items = self.data[split]
        if items and not os.path.isabs(items[0]["filepath"]):
            root = base_path if base_path is not None else os.path.dirname(os.path.abspath(data_path))
            for it in items:
                it["filepath"] = os.path.join(root, it["filepath"])
        items.sort(key=lambda x: (x["position"], int(os.path.basename(x["filepath"]).split(".")[0])))
        self._items = items
        self.data_pool = cycle(items) if cyclic else iter(items)

    @staticmethod
    def _load(path):
        with open(path, "rb") as fp:
            return pickle.load(fp, encoding="bytes")

    def get(self):
        try:
            ins = next(self.data_pool)
        except StopIteration:
            return None
        data = self._load(ins["filepath"])
        ee2base_pose = None
        if isinstance(data, dict):
            points, rgb, gt_pose = data["points"], data["rgb"], data["pose"]
            ee2base_pose = data.get("robot2ee_pose")
        else:
            points, rgb, _, _, gt_pose = data
        if gt_pose is not None:
            gt_pose = switch_w(gt_pose)
        if ee2base_pose is not None:
            ee2base_pose = switch_w(ee2base_pose)
        return PointCloud(points=points, rgb=rgb, ee2base_pose=ee2base_pose, timestamp=_utcnow(),
                          id=ins["filepath"], gt_pose=gt_pose)

    def run(self):
        return None

    def exit(self):
        return None

    def __len__(self):
        return len(self._items)


class PCDDataEngine:
    """app/data_engine.py:161-204. data_path: directory of <n>.pcd (+ <n>.npy gt pose, <n>_robot2ee_pose.npy), every
    `step`-th file in numeric order."""

    def __init__(self, data_path, cyclic=True, step=10):
        files = glob.glob(os.path.join(data_path, "*.pcd"))
        files.sort(key=lambda x: int(os.path.basename(x).split(".")[0]))
        self.data = [files[i] for i in range(0, len(files), step)]
        self.data_pool = cycle(self.data) if cyclic else iter(self.data)

    @staticmethod
    def _poses(path):
        gt_pose = ee2base = None
        p = path.replace(".pcd", ".npy")
        if os.path.exists(p):
            gt_pose = switch_w(np.load(p, allow_pickle=True))
        p = path.replace(".pcd", "_robot2ee_pose.npy")
        if os.path.exists(p):
            ee2base = switch_w(np.load(p, allow_pickle=True))
        return gt_pose, ee2base

    def get(self):
        try:
            ins = next(self.data_pool)
        except StopIteration:
            return None
        rec = pcd_records(ins)
        points = rec[:, :3].copy()
        rgb = unpack_rgb(rec[:, 3])
        m = _roi_mask(points)
        _, ee2base = self._poses(ins)
        return PointCloud(points=points[m], rgb=rgb[m], ee2base_pose=ee2base, timestamp=_utcnow(), id=ins,
                          gt_pose=None)

    def get_records(self):
        """the next frame as raw (x, y, z, packed rgb) records + its ee2base pose: input of ingest.ingest_clouds, which
        applies the finite / ROI filter and the colour unpacking on the device."""
        try:
            ins = next(self.data_pool)
        except StopIteration:
            return None
        _, ee2base = self._poses(ins)
        return pcd_records(ins), ee2base, ins

    def run(self):
        return None

    def exit(self):
        return None

    def __len__(self):
        return len(self.data)
