"""Generate tests/golden/reference_eval.npz: outputs of the reference's OWN evaluation metrics (utils/metrics.py:
compute_segmentation_metrics :50-107, compute_pose_metrics :110-127, compute_ADD_np :139-151), imported from
/root/reference in the authoring container, on seeded inputs.

    python tests/golden/make_golden_eval.py        # needs /root/reference
"""
import os
import sys
import types

import numpy as np

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    for name in ("ipdb",):
        m = types.ModuleType(name)
        m.set_trace = lambda *a, **k: None
        sys.modules[name] = m
    sys.path.insert(0, REF)
    from utils import metrics as RM
    rng = np.random.default_rng(21)
    out = {}
    segs = []
    for case in range(6):
        n = 4000
        gt = rng.integers(0, 3, n)
        pred = gt.copy()
        flip = rng.random(n) < (0.0, 0.02, 0.2, 0.5, 0.05, 0.0)[case]
        pred[flip] = rng.integers(0, 3, int(flip.sum()))
        if case == 4:
            gt[gt == 2] = 1          # a class absent from the ground truth (fn == 0 branch)
        if case == 5:
            pred[:] = gt             # perfect prediction (fp == 0 and fn == 0)
        r = RM.compute_segmentation_metrics(gt, pred)
        segs.append([r["accuracy"], r["precision"], r["recall"]] +
                    [r["class_results"][c][k] for c in ("background", "arm", "ee") for k in ("accuracy", "precision", "recall")])
        out[f"seg_gt{case}"], out[f"seg_pred{case}"] = gt.astype(np.int8), pred.astype(np.int8)
    out["seg_metrics"] = np.array(segs, dtype=np.float64)
    poses = rng.normal(size=(12, 7))
    poses[:, 3:] /= np.linalg.norm(poses[:, 3:], axis=1, keepdims=True)
    poses2 = poses + rng.normal(0, 0.05, poses.shape)
    poses2[:, 3:] /= np.linalg.norm(poses2[:, 3:], axis=1, keepdims=True)
    poses2[5, 3:] *= -1.0            # the negated quaternion is the same rotation
    pts = rng.normal(0, 0.05, (300, 3))
    pm = [RM.compute_pose_metrics(a.copy(), b.copy()) for a, b in zip(poses, poses2)]
    out.update(pose=poses, pose2=poses2, points=pts, dist=np.array([m["dist_position"] for m in pm]),
               angle=np.array([m["angle_diff"] for m in pm]),
               add=np.array([RM.compute_ADD_np(pts, a, b) for a, b in zip(poses, poses2)]))
    np.savez_compressed(os.path.join(HERE, "reference_eval.npz"), **out)
    print("seg metrics\n", np.round(out["seg_metrics"][:, :3], 4), "\nADD", np.round(out["add"][:4], 5))


if __name__ == "__main__":
    main()
