"""b200calib/evaluation.py (the JSON evaluation record that replaces app/test.py's xlsx sheet) against outputs of the
reference's own metric functions (tests/golden/make_golden_eval.py -> reference_eval.npz), and the aggregation rules of
app/test.py:239-286 (mean per position, then the mean of the position means)."""
import json
import os

import numpy as np

from conftest import GOLDEN
from b200calib import evaluation as E


def test_metrics_vs_reference_golden(built_lib):
    g = np.load(os.path.join(GOLDEN, "reference_eval.npz"))
    for case in range(len(g["seg_metrics"])):
        r = E.compute_segmentation_metrics(g[f"seg_gt{case}"], g[f"seg_pred{case}"])
        got = [r["accuracy"], r["precision"], r["recall"]] + \
              [r["class_results"][c][k] for c in E.CLASSES for k in ("accuracy", "precision", "recall")]
        assert np.allclose(got, g["seg_metrics"][case], rtol=0, atol=1e-12), case
    for i, (a, b) in enumerate(zip(g["pose"], g["pose2"])):
        m = E.compute_pose_metrics(a, b)
        assert abs(m["dist_position"] - g["dist"][i]) < 1e-12 and abs(m["angle_diff"] - g["angle"][i]) < 1e-10, i
        assert abs(E.compute_ADD_np(g["points"], a, b) - g["add"][i]) < 1e-12, i


def test_aggregation_and_report(tmp_path, built_lib):
    inst = [dict(position="p1", dist_position=dict(nn=0.01, nn_icp=0.004), angle_diff=dict(nn=0.1, nn_icp=0.05),
                 ADD_nn=0.02, mean_kp_error=0.03, is_confident=True),
            dict(position="p1", dist_position=dict(nn=0.03, nn_icp=0.006, kp=0.02), angle_diff=dict(nn=0.3, nn_icp=0.07, kp=0.2),
                 ADD_nn=0.04, is_confident=False),
            dict(position="p2", dist_position=dict(nn=0.05, nn_icp=0.01), angle_diff=dict(nn=0.5, nn_icp=0.09), ADD_nn=0.06,
                 is_confident=True)]
    rep = E.aggregate(inst, calibration_pose=np.array([0.1, 0.2, 0.3, 1, 0, 0, 0.0]),
                      gt_base2cam=np.array([0.1, 0.2, 0.31, 1, 0, 0, 0.0]))
    assert rep["positions"]["p1"]["dist_position_nn"] == [0.01, 0.03]
    assert abs(rep["overall"]["dist_position_nn"] - ((0.01 + 0.03) / 2 + 0.05) / 2) < 1e-15   # mean of position means
    assert abs(rep["overall"]["dist_position_kp"] - 0.02) < 1e-15                             # only where it exists
    assert abs(rep["overall"]["calibration_dist_position"] - 0.01) < 1e-12
    assert rep["frames"] == 3 and rep["frames_confident"] == 2
    path = E.write_report(str(tmp_path / "report.json"), rep)
    assert json.load(open(path))["overall"]["calibration_angle_diff"] == rep["overall"]["calibration_angle_diff"]
