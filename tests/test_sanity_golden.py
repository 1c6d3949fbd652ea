"""b200calib/sanity.py (host-side mirror of the reference's per-frame sanity check) against outputs of the
reference's OWN get_6_key_points / compute_kp_error / InferenceEngine.check_sanity (tests/golden/make_golden_sanity.py
-> reference_sanity.npz): key points to 1e-9 m, the index of the EE point under each key point and the verdict exact."""
import os

import numpy as np

from conftest import GOLDEN
from b200calib import sanity as S


def test_sanity_check_vs_reference_golden():
    g = np.load(os.path.join(GOLDEN, "reference_sanity.npz"))
    verdicts = []
    for c in range(len(g["sane"])):
        points, seg, pose = g["points"][c], g["seg"][c].astype(np.int64), g["pose"][c]
        ee = points[seg == 2]
        assert len(ee) == g["n_ee"][c]
        k1, i1 = S.get_6_key_points(ee, pose, switch_w=False, euclidean_threshold=0.04)
        assert (len(k1) == 0) == bool(g["empty_w"][c])
        if len(k1):
            assert np.abs(k1 - g["kp_w"][c]).max() < 1e-9, c
            assert np.array_equal(i1, g["idx_w"][c]), c
        pose_xyzw = np.concatenate((pose[:3], pose[4:], pose[3:4]))
        k2, i2 = S.get_6_key_points(ee, pose_xyzw)
        assert (len(k2) == 0) == bool(g["empty_x"][c])
        if len(k2):
            assert np.abs(k2 - g["kp_x"][c]).max() < 1e-9, c
            assert np.array_equal(i2, g["idx_x"][c]), c
        m = int(g["n_pred"][c])
        cls, xyz = g["pred_cls"][c][:m], g["pred_xyz"][c][:m]
        err = S.compute_kp_error(k1, xyz, cls)
        assert abs(err - g["kp_err"][c]) < 1e-9, c
        ok = S.check_sanity(points, seg, pose, [(int(a), b) for a, b in zip(cls, xyz)])
        assert ok == bool(g["sane"][c]), c
        verdicts.append(ok)
    assert 5 < sum(verdicts) < len(verdicts) - 5   # the fixture exercises both verdicts and every failure mode
