"""oracle/pipeline.py (the CPU restatement of InferenceEngine.predict) against outputs of the reference's OWN
InferenceEngine.predict run on the CPU (tests/golden/make_golden_predict.py -> reference_predict.npz): per-point labels
after the relabel + largest-cluster step bit-exact, EE pose (rotation network + "magic" translation) and base pose
(get_base2cam_pose) to 1e-6. The networks are rebuilt from the recorded seeds through the host-side mirror of the model
files (b200calib/models.py), whose initialisation reproduces the unchanged reference classes (checked by weight sums)."""
import os

import numpy as np
import torch

from conftest import GOLDEN
import oracle.MinkowskiEngine as OME
from oracle import pipeline as OP
from b200calib.models import make_models, randomize_bn_stats


def _nets(g):
    M = make_models(OME)
    torch.manual_seed(int(g["seed_seg"]))
    seg = randomize_bn_stats(M.RobotNetSegmentation(3, num_classes=3), int(g["seed_seg"])).eval()
    with torch.no_grad():
        seg.regression[2].linear.bias.copy_(torch.from_numpy(g["seg_head_bias"]))
    torch.manual_seed(int(g["seed_rot"]))
    rot = randomize_bn_stats(M.RobotNetEncode(3, 7), int(g["seed_rot"])).eval()
    return seg, rot


def test_oracle_pipeline_vs_reference_predict_golden():
    g = np.load(os.path.join(GOLDEN, "reference_predict.npz"))
    seg, rot = _nets(g)
    ws = float(sum(v.double().abs().sum() for v in seg.state_dict().values()))
    wr = float(sum(v.double().abs().sum() for v in rot.state_dict().values()))
    assert abs(ws - float(g["seg_weight_sum"])) < 1e-6 * ws, "mirror initialisation differs from the reference classes"
    assert abs(wr - float(g["rot_weight_sum"])) < 1e-6 * wr
    cfg = dict(seg_scale=float(g["seg_scale"]), rot_scale=float(g["rot_scale"]),
               ee_point_counts_threshold=int(g["ee_threshold"]), icp_enabled=False)
    for i in range(2):
        r = OP.predict_frame(dict(seg=seg, rot=rot, kp=None), None, g[f"f{i}_points"], g[f"f{i}_rgb255"], cfg,
                             ee2base_pose=g["ee2base"])
        assert np.array_equal(r["segmentation"], g[f"f{i}_segmentation"].astype(np.int64)), f"frame {i}: labels"
        assert (r["segmentation"] == 2).sum() > 512 and (r["segmentation"] == 1).sum() > 0
        ee, base = g[f"f{i}_ee_pose"], g[f"f{i}_base_pose"]
        assert np.allclose(r["ee_pose"][:3], ee[:3], atol=1e-6) and np.allclose(r["ee_pose"][3:], ee[3:], atol=1e-6)
        assert np.allclose(r["base_pose"][:3], base[:3], atol=1e-6)
        assert min(np.abs(r["base_pose"][3:] - base[3:]).max(), np.abs(r["base_pose"][3:] + base[3:]).max()) < 1e-6


def test_mirror_models_reproduce_unchanged_reference_logits():
    """b200calib/models.py (the mirror the GPU box runs) against the UNCHANGED reference classes on the same substrate
    (the oracle package): voxel coordinates and voxel logits of RobotNetSegmentation and the raw RobotNetEncode output
    recorded by make_golden_predict.py are reproduced to fp32 round-off - the mirror IS the reference topology."""
    g = np.load(os.path.join(GOLDEN, "reference_predict.npz"))
    seg, rot = _nets(g)
    for i in range(2):
        pts = torch.from_numpy(g[f"f{i}_points"])
        rgbn = torch.from_numpy(g[f"f{i}_rgb255"] / 255.0 - 0.5).to(torch.float32)
        fld = OME.TensorField(features=rgbn, coordinates=OME.utils.batched_coordinates([pts * float(g["seg_scale"])],
                                                                                     dtype=torch.float32))
        with torch.no_grad():
            out = seg(fld.sparse())
        assert np.array_equal(out.C.numpy(), g[f"f{i}_voxel_coords"])
        assert np.allclose(out.F.numpy(), g[f"f{i}_voxel_logits"], rtol=0, atol=1e-6)
        assert np.array_equal(out.slice(fld).F.max(1)[1].numpy(), g[f"f{i}_point_labels_raw"])
