"""Generate tests/golden/reference_predict.npz by running the reference's OWN InferenceEngine.predict
(app/inference_engine.py:281-382: predict_segmentation, predict_rotation, predict_translation, get_base2cam_pose) on
the CPU, with this repo's oracle package standing in for MinkowskiEngine, the unchanged reference model files
(random init under fixed seeds; the checkpoints are unpublished) and stubs for what is not installed.

What is patched, and why:
  * ipdb / open3d / tensorboardX / openpyxl: stub modules (not installed; not touched by the code that runs)
  * cluster_util: scikit-learn 1.9 rejects the `affinity=` keyword of utils/output.py:15-20, so the same
    AgglomerativeClustering is built with `metric=` (the renamed keyword)
  * ICP off (Open3D absent); key points off (num_of_dense_input_points raised: np.random sampling and PointNet++ are
    pinned separately in reference_pointnet2.npz); check_sanity returns at its first test (min_num_of_ee_points raised)

    python tests/golden/make_golden_predict.py        # needs /root/reference
"""
import os
import sys
import tempfile
import types
from datetime import datetime

import numpy as np
import torch

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
SEED_SEG, SEED_ROT = 101, 102


def main():
    tmp = tempfile.mkdtemp(prefix="b2me_golden_")
    sys.argv = ["x", "--config", os.path.join(REF, "config", "default.yaml"), "--log_path", os.path.join(tmp, "log.log"),
                "--exp_path", os.path.join(tmp, "exp")]
    for name in ("ipdb", "open3d", "tensorboardX", "openpyxl", "turtle"):
        m = types.ModuleType(name)
        m.set_trace = lambda *a, **k: None
        m.SummaryWriter = object
        m.pos = None
        sys.modules[name] = m
    if not hasattr(np, "int"):
        np.int = int      # NumPy 2 removed the alias the reference uses (app/inference_engine.py:283)
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "markerless-robot-camera-calibration_b200"))
    import oracle.MinkowskiEngine as OME
    import oracle.MinkowskiEngine.modules.resnet_block as rb
    import oracle.MinkowskiEngine.utils as mu
    import oracle.MinkowskiEngine.MinkowskiOps as mo
    sys.modules["MinkowskiEngine"] = OME
    sys.modules["MinkowskiEngine.modules"] = OME.modules
    sys.modules["MinkowskiEngine.modules.resnet_block"] = rb
    sys.modules["MinkowskiEngine.utils"] = mu
    sys.modules["MinkowskiEngine.MinkowskiOps"] = mo
    sys.path.insert(0, REF)
    sys.path.insert(0, os.path.join(REF, "app"))
    from app import inference_engine as IE
    from dto import PointCloudDTO
    from model.robotnet_segmentation import RobotNetSegmentation
    from model.robotnet_encode import RobotNetEncode
    from sklearn.cluster import AgglomerativeClustering
    from b200calib.models import randomize_bn_stats
    from b200calib.synthetic import make_frame

    torch.manual_seed(SEED_SEG)
    seg = randomize_bn_stats(RobotNetSegmentation(3, num_classes=3), SEED_SEG).eval()
    # the random-init head prefers one class everywhere; tilt its bias so that ~35 % of the points of a probe frame are
    # labelled EE (class 2) and ~30 % arm: the EE gate, the relabel + largest-cluster step and the pose stages then run
    probe = make_frame(769, width=128, height=96)
    with torch.no_grad():
        pp = torch.from_numpy(probe["points"])
        fld = OME.TensorField(features=torch.from_numpy(probe["rgb"]) - 0.5,
                              coordinates=OME.utils.batched_coordinates([pp * 200.0], dtype=torch.float32))
        lg = seg(fld.sparse()).slice(fld).F
        b2 = torch.quantile(torch.maximum(lg[:, 0], lg[:, 1]) - lg[:, 2], 0.35)
        b1 = torch.quantile(lg[:, 0] - lg[:, 1], 0.45)
        seg.regression[2].linear.bias.add_(torch.tensor([0.0, float(b1), float(b2)]))
        out_bias = seg.regression[2].linear.bias.detach().clone().numpy()
    torch.manual_seed(SEED_ROT)
    rot = randomize_bn_stats(RobotNetEncode(3, 7), SEED_ROT).eval()

    class _Cluster:  # utils/output.py:13-28 with the renamed sklearn keyword
        def __init__(self):
            self.cluster = AgglomerativeClustering(distance_threshold=0.06, n_clusters=None, metric="euclidean",
                                                   linkage="single")
        get_largest_cluster = IE.out_utils.ClusterUtil.get_largest_cluster

    cfg = IE._config
    cfg.INFERENCE.icp_enabled = False
    cfg.INFERENCE.num_of_dense_input_points = 10 ** 9
    cfg.INFERENCE.SANITY.min_num_of_ee_points = 10 ** 9
    eng = object.__new__(IE.InferenceEngine)
    eng.pred_enabled = True
    eng._segmentation_model, eng._rotation_model, eng._key_points_model = seg, rot, None
    eng.cluster_util = _Cluster()
    eng.reference_key_points = np.zeros((6, 3))
    eng.camera_link_transformation_pose = None

    out = dict(seed_seg=SEED_SEG, seed_rot=SEED_ROT, seg_head_bias=out_bias,
               seg_weight_sum=float(sum(v.double().abs().sum() for v in seg.state_dict().values())),
               rot_weight_sum=float(sum(v.double().abs().sum() for v in rot.state_dict().values())),
               seg_scale=float(cfg.INFERENCE.SEGMENTATION.scale), rot_scale=float(cfg.INFERENCE.ROTATION.scale),
               ee_threshold=int(cfg.INFERENCE.ee_point_counts_threshold))
    ee2base = np.array([0.4, -0.1, 0.3, 0.9238795, 0.0, 0.3826834, 0.0])
    for i, (w, h) in enumerate(((128, 96), (160, 120))):
        f = make_frame(770 + i, width=w, height=h)
        rgb255 = np.round(f["rgb"] * 255.0).astype(np.float32)       # a camera's 0..255 colours: exercises /255
        data = PointCloudDTO(points=f["points"], rgb=rgb255, timestamp=datetime.utcnow(), ee2base_pose=ee2base)
        r = eng.predict(data)
        n_ee = int((r.segmentation == 2).sum())
        print(f"frame {i}: {len(f['points'])} points, {n_ee} EE points after the cluster filter, ee_pose",
              None if r.ee_pose is None else np.round(r.ee_pose, 4))
        # voxel logits of the UNCHANGED reference RobotNetSegmentation on this frame, through the reference's own call
        # pattern (app/inference_engine.py:405-417), and the raw 7-vector of the unchanged RobotNetEncode on the EE crop
        # (app/inference_engine.py:437-457): what the CUDA package has to reproduce through the mirror model files
        with torch.no_grad():
            rgbn = torch.from_numpy(rgb255 / 255.0 - 0.5).to(torch.float32)
            pts_t = torch.from_numpy(f["points"])
            fld = OME.TensorField(features=rgbn, quantization_mode=OME.SparseTensorQuantizationMode.UNWEIGHTED_AVERAGE,
                                  minkowski_algorithm=OME.MinkowskiAlgorithm.SPEED_OPTIMIZED,
                                  coordinates=OME.utils.batched_coordinates([pts_t * cfg.INFERENCE.SEGMENTATION.scale],
                                                                            dtype=torch.float32))
            sout = seg(fld.sparse())
            out[f"f{i}_voxel_coords"] = sout.C.numpy().astype(np.int32)
            out[f"f{i}_voxel_logits"] = sout.F.numpy().astype(np.float32)
            out[f"f{i}_point_labels_raw"] = sout.slice(fld).F.max(1)[1].numpy().astype(np.int8)
            ee = np.where(r.segmentation == 2)[0]
            ee_pts = f["points"][ee]
            ee_pts = ee_pts - (ee_pts.max(0) + ee_pts.min(0)) / 2          # utils/preprocess.py:8-11
            rfld = OME.TensorField(features=rgbn[ee], quantization_mode=OME.SparseTensorQuantizationMode.UNWEIGHTED_AVERAGE,
                                   minkowski_algorithm=OME.MinkowskiAlgorithm.SPEED_OPTIMIZED,
                                   coordinates=OME.utils.batched_coordinates(
                                       [torch.from_numpy(ee_pts) * cfg.INFERENCE.ROTATION.scale], dtype=torch.float32))
            out[f"f{i}_rot_out"] = rot(rfld.sparse())[0].numpy().astype(np.float32)
        out[f"f{i}_points"] = f["points"]
        out[f"f{i}_rgb255"] = rgb255
        out[f"f{i}_segmentation"] = r.segmentation.astype(np.int8)
        out[f"f{i}_ee_pose"] = np.full(7, np.nan) if r.ee_pose is None else np.asarray(r.ee_pose, dtype=np.float64)
        out[f"f{i}_base_pose"] = np.full(7, np.nan) if r.base_pose is None else np.asarray(r.base_pose, dtype=np.float64)
    out["ee2base"] = ee2base
    np.savez_compressed(os.path.join(HERE, "reference_predict.npz"), **out)


if __name__ == "__main__":
    main()
