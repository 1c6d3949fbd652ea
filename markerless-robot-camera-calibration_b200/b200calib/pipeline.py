"""Batched mirror of app/inference_engine.py InferenceEngine.predict (:281-382) on one GPU.

The reference runs one frame at a time with >= 4 device->host syncs and CPU stages (sklearn clustering,
Open3D ICP) in between. Here a batch of frames goes through every stage on the device:

  normalise colours (utils/preprocess.py:20-37)             torch elementwise (plumbing)
  predict_segmentation (:395-435)                           K1 voxelise -> MinkUNet (K2-K4) -> head (K5) ->
                                                            per-point labels (fused slice+argmax) -> K7 largest
                                                            EE cluster per frame
  gate on ee_point_counts_threshold (:295)                  host, from the per-frame EE counts
  predict_rotation (:437-457)                               K1 + RobotNetEncode trunk (K2-K4, K6) on the EE crops
  predict_translation "magic" (:459-489)                    fused min/max reduction kernel
  predict_key_points, ME branch (:539-555)                  K1 + key-point seg-net + K8a
  predict_pose_from_kp (:384-393)                           K9 batched Kabsch
  match_icp x2 (:358-362, utils/icp.py:50-81)               K10 batched ICP (+K9)
  get_base2cam_pose (:366-369)                              host 4x4 algebra

Frames are independent, so multi-GPU runs shard frames across ranks (dist.py) and all-gather the poses.
"""
import time
from dataclasses import dataclass, field
from typing import List, Optional

import numpy as np
import torch

import MinkowskiEngine as ME
from MinkowskiEngine._lib import lib, check, ptr, stream
from MinkowskiEngine.core import _count

from . import output as out_utils
from .icp import icp_p2p_batched
from .synthetic import REFERENCE_KEY_POINTS
from .transformation import (rigid_transform_3D_batched, get_poses_from_matrices, get_base2cam_pose)


@dataclass
class PipelineConfig:
    """thresholds and scales of config/default.yaml:136-191 (INFERENCE section)."""
    seg_scale: float = 200.0
    rot_scale: float = 200.0
    kp_scale: float = 800.0
    ee_point_counts_threshold: int = 512
    cluster_dist: float = 0.06
    kp_conf_threshold: float = 0.75
    icp_enabled: bool = True
    rot_center_at_origin: bool = True
    kp_center_at_origin: bool = True
    translation_x_offset: float = -0.015
    num_dense_points: int = 2048   # INFERENCE.num_of_dense_input_points (PointNet++ key-point branch)
    # is_confident = check_sanity(...) exactly as the reference sets it (app/inference_engine.py:323), evaluated on the
    # device for every posed frame (b2me_sanity_check). Off: no frame is ever marked confident, so the calibration tail
    # (calibration.calibrate_individual filters on is_confident) cannot silently average unchecked frames.
    sanity_check: bool = True
    vote_ee_r: float = 0.02               # PARAM.ee_r (config/default.yaml:8): offset of the voted centre along -x of the EE
    sanity_min_ee_points: int = 2048      # INFERENCE.SANITY.min_num_of_ee_points
    sanity_kp_error_margin: float = 0.05  # INFERENCE.KEY_POINTS.error_margin


@dataclass
class FrameResult:
    """ResultDTO of app/dto.py:37-47 (poses are x,y,z,qw,qx,qy,qz)."""
    segmentation: np.ndarray
    ee_pose: Optional[np.ndarray] = None
    base_pose: Optional[np.ndarray] = None
    key_points: list = field(default_factory=list)
    key_points_pose: Optional[np.ndarray] = None
    key_points_base_pose: Optional[np.ndarray] = None
    is_confident: bool = False
    icp_stats: Optional[np.ndarray] = None
    vote_center: Optional[np.ndarray] = None   # get_pred_center of the vote head (utils/output.py:45-64), when it runs


def normalize_colors_(rgb, bidx=None, offs=None):
    """utils/preprocess.py:20-37 per frame on the device (b2me_normalize_colors: / 255, min-max branch for negative
    inputs, - 0.5; every decision from the frame's own extrema, no host round trip). Without bidx / offs the tensor is
    one frame."""
    n = rgb.shape[0]
    if n == 0:
        return rgb
    if bidx is None:
        bidx = torch.zeros((n,), dtype=torch.float32, device=rgb.device)
        offs = np.array([0, n], dtype=np.int32)
    return out_utils.normalize_colors_batched(rgb, bidx, offs)


def batch_frames(frames, device, pinned=None):
    """list of (points [N_i,3] f32, rgb [N_i,3] f32) host arrays -> device tensors + frame offsets.
    One H2D copy per tensor per batch (from pinned staging when given)."""
    counts = [len(f[0]) for f in frames]
    offs = np.zeros(len(frames) + 1, dtype=np.int64)
    np.cumsum(counts, out=offs[1:])
    N = int(offs[-1])
    if pinned is None:
        pts = torch.empty((N, 3), dtype=torch.float32).pin_memory()
        rgb = torch.empty((N, 3), dtype=torch.float32).pin_memory()
        bidx = torch.empty((N,), dtype=torch.float32).pin_memory()
    else:
        pts, rgb, bidx = pinned[0][:N], pinned[1][:N], pinned[2][:N]
    for i, f in enumerate(frames):
        pts[offs[i]:offs[i + 1]] = torch.as_tensor(f[0])
        rgb[offs[i]:offs[i + 1]] = torch.as_tensor(f[1])
        bidx[offs[i]:offs[i + 1]] = float(i)
    return (pts.to(device, non_blocking=True), rgb.to(device, non_blocking=True),
            bidx.to(device, non_blocking=True), offs)


def _field(points, rgb, bidx, scale, nb):
    coords = torch.cat((bidx.unsqueeze(1), points * scale), dim=1)  # fp32 multiply, then K1 floors (§8a a2)
    f = ME.TensorField(features=rgb, coordinates=coords,
                       quantization_mode=ME.SparseTensorQuantizationMode.UNWEIGHTED_AVERAGE,
                       minkowski_algorithm=ME.MinkowskiAlgorithm.SPEED_OPTIMIZED, device=points.device)
    f._nb = nb
    return f


def segment_points(model, points, rgb, bidx, nb, scale):
    """predict_segmentation up to the per-point arg-max: returns labels [N] uint8 (device)."""
    fld = _field(points, rgb, bidx, scale, nb)
    sp = fld.sparse()
    out = model(sp)
    logits = out.F  # [V, C] f32 voxel logits (K5)
    V, C = logits.shape
    # slice + arg-max fused: arg-max per voxel (by-product of the head's last linear), then 1 byte per point
    # through the inverse map
    vlab = getattr(out, "_row_argmax", None)
    if vlab is None:
        vlab = torch.empty((max(V, 1),), dtype=torch.uint8, device=logits.device)
        check(lib.b2me_linear_small(ptr(logits), 0, V, C, ptr(_eye(C, logits.device)), None, C, None, ptr(vlab),
                                    stream()), "argmax")
        _count(1)
    N = points.shape[0]
    labels = torch.empty((max(N, 1),), dtype=torch.uint8, device=logits.device)
    check(lib.b2me_gather_labels(ptr(vlab), ptr(fld.inverse_mapping), N, ptr(labels), stream()), "gather_labels")
    _count(1)
    return labels[:N], fld, out


_EYES = {}


def _eye(C, device):
    k = (C, str(device))
    if k not in _EYES:
        _EYES[k] = torch.eye(C, dtype=torch.float32, device=device).contiguous()
    return _EYES[k]


class _Stage:
    """opt-in per-stage wall times (device-synchronised on both sides); used by bench.py --stages only."""

    def __init__(self, eng, name):
        self.eng, self.name = eng, name

    def __enter__(self):
        if self.eng.stage_times is not None:
            torch.cuda.synchronize()
            self.t0 = time.perf_counter()

    def __exit__(self, *a):
        if self.eng.stage_times is not None:
            torch.cuda.synchronize()
            d = self.eng.stage_times
            d[self.name] = d.get(self.name, 0.0) + (time.perf_counter() - self.t0) * 1e3


class BatchedInferenceEngine:
    stage_times = None  # dict name -> ms when enabled

    def __init__(self, seg_model, rot_model=None, kp_model=None, cad_points=None, config=None,
                 reference_key_points=REFERENCE_KEY_POINTS, vote_model=None):
        self.vote_model = vote_model.eval() if vote_model is not None else None   # RobotNetVote (model/robotnet_vote.py)
        self.seg_model = seg_model.eval()
        self.rot_model = rot_model.eval() if rot_model is not None else None
        self.kp_model = kp_model.eval() if kp_model is not None else None
        self.cad = cad_points
        self.cfg = config or PipelineConfig()
        self.ref_kp = torch.as_tensor(np.asarray(reference_key_points, dtype=np.float64))

    @torch.no_grad()
    def segment(self, points, rgb, bidx, offs):
        """labels after the EE largest-cluster filter (app/inference_engine.py:419-433), kept EE rows, per-frame
        offsets of those rows (host)."""
        nb = len(offs) - 1
        offs_dev = torch.as_tensor(np.asarray(offs), dtype=torch.int32).to(points.device)
        labels, _, _ = segment_points(self.seg_model, points, normalize_colors_(rgb, bidx, offs_dev), bidx, nb,
                                      self.cfg.seg_scale)
        return self.filter_ee(points, labels, offs_dev, nb)

    def filter_ee(self, points, labels, offs_dev, nb):
        """app/inference_engine.py:422-433 for a batch: every class-2 point becomes class 1, the largest single-linkage
        cluster (K7) of each frame's class-2 points becomes class 2 again. Two order-preserving selections
        (b2me_select_rows) + K7, two host syncs (EE counts, kept counts). Returns (labels [N] u8, kept rows [m] i32
        frame-sorted, per-frame offsets of the kept rows [nb+1] host int64)."""
        dev = points.device
        N = labels.shape[0]
        zero_offs = torch.zeros(nb + 1, dtype=torch.int64)
        if N == 0:
            return labels, torch.zeros((0,), dtype=torch.int32, device=dev), zero_offs
        ws = torch.empty((lib.b2me_select_workspace_bytes(N),), dtype=torch.uint8, device=dev)
        ee_rows = torch.empty((N,), dtype=torch.int32, device=dev)
        ee_offs_d = torch.empty((nb + 1,), dtype=torch.int32, device=dev)
        check(lib.b2me_select_rows(ptr(labels), 2, None, N, ptr(offs_dev), nb, ptr(ee_rows), ptr(ee_offs_d), ptr(ws),
                                   ws.numel(), stream()), "select_rows")
        _count(5)
        labels2 = torch.where(labels == 2, torch.ones_like(labels), labels)
        ee_offs_h = ee_offs_d.cpu().numpy()                        # host sync 1: EE points per frame
        n_ee = int(ee_offs_h[-1])
        if n_ee == 0:
            return labels2, ee_rows[:0], zero_offs
        ee_pts = torch.empty((n_ee, 3), dtype=torch.float32, device=dev)
        check(lib.b2me_gather_crops(ptr(points), None, ptr(ee_rows), n_ee, None, 0, ptr(ee_pts), None, None, stream()),
              "gather_crops")
        _count(1)
        mask, sizes = out_utils.largest_cluster_mask(ee_pts, ee_offs_h.astype(np.int32), self.cfg.cluster_dist)
        keep = torch.empty((n_ee,), dtype=torch.int32, device=dev)
        koffs_d = torch.empty((nb + 1,), dtype=torch.int32, device=dev)
        check(lib.b2me_select_rows(ptr(mask), -1, ptr(ee_rows), n_ee, ptr(ee_offs_d), nb, ptr(keep), ptr(koffs_d),
                                   ptr(ws), ws.numel(), stream()), "select_rows")
        _count(5)
        host = torch.cat((koffs_d, sizes)).cpu().numpy()           # host sync 2: kept points per frame + K7 status
        if int(host[nb + 1:].min()) < 0:
            raise RuntimeError("largest-cluster filter: a non-finite or far out-of-range EE point (b2me_largest_cluster)")
        koffs = torch.from_numpy(host[:nb + 1].astype(np.int64))
        keep = keep[:int(koffs[-1])]
        labels2[keep.long()] = 2
        return labels2, keep, koffs

    @torch.no_grad()
    def pose_from_ee(self, points, rgb_norm, ee_idx, ee_offs, ee2base_poses=None, kp_conf_threshold=None):
        """rotation + translation + key points + Kabsch + sanity check + ICP for the frames whose EE crop passes the
        gate. ee_idx: rows of the EE points (frame-sorted, i32), ee_offs [nb+1] host int64."""
        cfg = self.cfg
        nb = len(ee_offs) - 1
        dev = points.device
        counts = (ee_offs[1:] - ee_offs[:-1]).numpy()
        ok = np.nonzero(counts >= cfg.ee_point_counts_threshold)[0]
        res = dict(ok_frames=ok, ee_pose=None, kp_pose=None, icp_stats=None, key_points=None)
        if len(ok) == 0 or self.rot_model is None:
            return res
        # compact the crops of the frames that pass the gate (all of them, usually: no copy of the row list then)
        S = len(ok)
        if S == nb:
            sel = ee_idx
        else:
            sel = torch.cat([ee_idx[int(ee_offs[f]):int(ee_offs[f + 1])] for f in ok])
        soffs = np.zeros(S + 1, dtype=np.int32)
        np.cumsum(counts[ok], out=soffs[1:])
        soffs_dev = torch.from_numpy(soffs).to(dev)
        m = int(soffs[-1])
        pts = torch.empty((m, 3), dtype=torch.float32, device=dev)
        feats = torch.empty((m, 3), dtype=torch.float32, device=dev)
        segf = torch.empty((m,), dtype=torch.float32, device=dev)
        check(lib.b2me_gather_crops(ptr(points), ptr(rgb_norm), ptr(sel.contiguous()), m, ptr(soffs_dev), S, ptr(pts),
                                    ptr(feats), ptr(segf), stream()), "gather_crops")
        # center_at_origin of every crop (utils/preprocess.py:8-11), shared by the rotation and key-point networks
        center = torch.empty((S, 3), dtype=torch.float32, device=dev)
        centered = torch.empty((m, 3), dtype=torch.float32, device=dev)
        check(lib.b2me_center_segments(ptr(pts), ptr(soffs_dev), S, ptr(center), ptr(centered), stream()),
              "center_segments")
        _count(2)

        # --- rotation (app/inference_engine.py:437-457)
        with _Stage(self, "rotation"):
            rot_pts = centered if cfg.rot_center_at_origin else pts
            fld = _field(rot_pts, feats, segf, cfg.rot_scale, S)
            rot_out = self.rot_model(fld.sparse())          # [S, 7|10]
            quat = rot_out[:, 3:7].float().contiguous()     # W,X,Y,Z
        # --- translation (app/inference_engine.py:459-489)
        with _Stage(self, "translation"):
            pos = out_utils.translation_magic_batched(pts, soffs, quat, cfg.translation_x_offset)
            ee_pose = torch.cat((pos, quat.double()), dim=1)  # x,y,z,qw,qx,qy,qz

        # --- vote head (model/robotnet_vote.py:36-71 + utils/output.py:45-64, the path of test_vote.py:75-101): per-point
        #     logits of RobotNetVote on the crop, mean coordinate of the 8 points with the largest "centre" logit, moved
        #     by R(q) [-ee_r, 0, 0] with the rotation network's quaternion. Optional: InferenceEngine.predict does not
        #     call it.
        vote_center = None
        if self.vote_model is not None:
            with _Stage(self, "vote"):
                vfld = _field(rot_pts, feats, segf, cfg.rot_scale, S)
                vout = self.vote_model(vfld.sparse()).slice(vfld).F.float().contiguous()
                ctr = out_utils.vote_centers_batched(vout, pts, soffs, col=1, topk=8)          # [S,3] f32
                off = torch.tensor([-cfg.vote_ee_r, 0.0, 0.0], dtype=torch.float32, device=dev)
                qn = quat / quat.norm(dim=1, keepdim=True)
                Rq = _poses_to_matrices(torch.cat((torch.zeros((S, 3), dtype=torch.float64, device=dev), qn.double()),
                                                  dim=1))[:, :3, :3].float()
                vote_center = ctr + (Rq @ off)

        # --- key points (app/inference_engine.py:491-559) + Kabsch (:384-393)
        kp_T = None
        bp = kp_xyz = None
        th = cfg.kp_conf_threshold if kp_conf_threshold is None else kp_conf_threshold
        if self.kp_model is not None:
            with _Stage(self, "key_points"):
                kp_pts = centered if cfg.kp_center_at_origin else pts
                if getattr(self.kp_model, "is_dense_pointnet2", False) or type(self.kp_model).__name__ == "PointNet2SSG":
                    # the reference's default branch (:511-537): PointNet++ on num_of_dense_input_points points drawn
                    # uniformly without replacement from every EE crop, here for all crops in one batch
                    bp, bi = self._key_points_pointnet2(kp_pts, feats, segf.long(), soffs, S)
                else:
                    # MinkUNet branch (:539-555)
                    kfld = _field(kp_pts, feats, segf, cfg.kp_scale, S)
                    kout = self.kp_model(kfld.sparse()).slice(kfld).F.float()
                    bp, bi = out_utils.key_point_predictions_batched(kout, soffs)
                K = bp.shape[1]
                valid = bp > th                                     # [S,K]
                nvalid = valid.sum(1)
                # pack the selected (reference kp, predicted point) pairs to the front of each row
                order = torch.argsort((~valid).to(torch.int8), dim=1, stable=True)
                ref = self.ref_kp.to(dev)[: K][order]               # [S,K,3]
                kp_xyz = pts[bi.long().clamp(min=0)]                # [S,K,3] f32: the EE point under every key point
                tgt = torch.gather(kp_xyz.double(), 1, order.unsqueeze(-1).expand(-1, -1, 3))
                R, t = rigid_transform_3D_batched(ref.contiguous(), tgt.contiguous(), nvalid.to(torch.int32))
                kp_T = torch.zeros((S, 4, 4), dtype=torch.float64, device=dev)
                kp_T[:, :3, :3], kp_T[:, :3, 3], kp_T[:, 3, 3] = R, t, 1.0
                res["key_points"] = (bp, bi - torch.as_tensor(soffs[:-1], device=dev).unsqueeze(1), nvalid)
                kp_ok = nvalid >= 4

        # --- is_confident = check_sanity(data, result) (app/inference_engine.py:323, :246-279): the EE pose BEFORE the
        #     ICP refinement against the corners / gripper tips found on the crop, and the selected key points
        confident = None
        if cfg.sanity_check:
            with _Stage(self, "sanity"):
                K6 = min(bp.shape[1], 6) if bp is not None else 0
                confident = out_utils.sanity_check_batched(
                    pts, soffs, ee_pose, bp[:, :K6].contiguous() if K6 else None,
                    kp_xyz[:, :K6].contiguous() if K6 else None, th, cfg.sanity_min_ee_points,
                    cfg.sanity_kp_error_margin)

        # --- ICP refinement of both poses (app/inference_engine.py:358-362): ONE launch, 2S problems
        with _Stage(self, "icp"):
            ee_T = _poses_to_matrices(ee_pose)
            kp_T0 = kp_T          # the key-point pose before its ICP refinement (evaluation: "kp" vs "kp_icp")
            stats = kstats = None
            if cfg.icp_enabled and self.cad is not None:
                if kp_T is not None:
                    both_T = torch.cat((ee_T, kp_T))
                    both_pts = torch.cat((pts, pts))
                    both_offs = np.concatenate((soffs, soffs[1:] + soffs[-1])).astype(np.int32)
                    T_all, st = icp_p2p_batched(self.cad, both_pts, both_offs, both_T)
                    ee_T, kp_T, stats, kstats = T_all[:S], T_all[S:], st[:S], st[S:]
                else:
                    ee_T, stats = icp_p2p_batched(self.cad, pts, soffs, ee_T)
        # --- one device->host transfer of everything the host needs
        with _Stage(self, "readback"):
            conf_col = (confident.double() if confident is not None
                        else torch.zeros((S,), dtype=torch.float64, device=dev)).unsqueeze(1)
            vote_col = (vote_center.double() if vote_center is not None
                        else torch.full((S, 3), float("nan"), dtype=torch.float64, device=dev))
            pack = [ee_T.reshape(S, 16), ee_pose, conf_col, vote_col]   # ee_pose: before the ICP refinement
            if stats is not None:
                pack.append(stats)
            if kp_T is not None:
                pack += [kp_T.reshape(S, 16), kp_ok.double().unsqueeze(1), res["key_points"][0].double(),
                         res["key_points"][1].double(), res["key_points"][2].double().unsqueeze(1)]
                if kstats is not None:
                    pack.append(kstats)
                pack.append(kp_T0.reshape(S, 16))
            host = torch.cat(pack, dim=1).cpu().numpy()
            c = 27
            res["ee_T"] = host[:, :16].reshape(S, 4, 4)
            res["ee_pose_initial"] = host[:, 16:23]
            res["confident"] = host[:, 23] > 0.5
            res["vote_center"] = host[:, 24:27] if vote_center is not None else None
            res["kp_threshold"] = th
            if stats is not None:
                res["icp_stats"] = host[:, c:c + 4]
                c += 4
            res["ee_pose"] = get_poses_from_matrices(res["ee_T"])
            if kp_T is not None:
                res["kp_T"] = host[:, c:c + 16].reshape(S, 4, 4)
                res["kp_valid"] = host[:, c + 16] > 0.5
                c += 17
                K = bp.shape[1]
                res["key_points"] = (host[:, c:c + K].astype(np.float32), host[:, c + K:c + 2 * K].astype(np.int64),
                                     host[:, c + 2 * K].astype(np.int64))
                c += 2 * K + 1
                if kstats is not None:
                    res["kp_icp_stats"] = host[:, c:c + 4]
                    c += 4
                res["kp_pose"] = get_poses_from_matrices(res["kp_T"])
                res["kp_pose_initial"] = get_poses_from_matrices(host[:, c:c + 16].reshape(S, 4, 4))
        return res

    def _key_points_pointnet2(self, kp_pts, feats, seg_ids, soffs, S):
        """PointNet2SSG key-point logits for S crops -> (best_prob [S,K], best_idx [S,K] rows of the compacted crops).
        Every crop contributes `cfg.num_dense_points` points (uniform sample without replacement, the reference's
        np.random.choice at app/inference_engine.py:514-518); crops with fewer points than that get no key points
        (the reference returns [] at :512-513)."""
        nd = self.cfg.num_dense_points
        dev = kp_pts.device
        counts = torch.as_tensor(np.diff(soffs), device=dev)
        starts = torch.as_tensor(soffs[:-1].astype(np.int64), device=dev)
        # random order inside every crop: sort by (crop, uniform key); the first nd rows of a crop are its sample
        key = seg_ids.double() + torch.rand(kp_pts.shape[0], device=dev, dtype=torch.float64)
        order = torch.argsort(key)
        K = self.kp_model.conv2.out_channels
        bp = torch.zeros((S, K), dtype=torch.float32, device=dev)
        bi = torch.full((S, K), -1, dtype=torch.int32, device=dev)
        big = torch.nonzero(counts >= nd).flatten()
        if big.numel() == 0:
            return bp, bi
        take = (starts[big].unsqueeze(1) + torch.arange(nd, device=dev).unsqueeze(0)).reshape(-1)
        sample = order[take]                                            # [nb * nd] rows of the compacted crops
        inp = torch.cat((kp_pts[sample], feats[sample]), dim=1).view(big.numel(), nd, -1).transpose(2, 1).contiguous()
        logits = self.kp_model(inp)[0].reshape(big.numel() * nd, -1).float()
        sub_offs = np.arange(big.numel() + 1, dtype=np.int32) * nd
        p, i = out_utils.key_point_predictions_batched(logits.contiguous(), sub_offs)
        bp[big] = p
        bi[big] = sample[i.long()].to(torch.int32)
        self.last_kp_sample = (big, sample.view(big.numel(), nd))
        return bp, bi

    @torch.no_grad()
    def predict_device(self, points, rgb, bidx, offs, ee2base_poses=None, gt_labels=None, kp_conf_threshold=None,
                       rgb_normalized=False):
        """the whole per-batch pipeline on tensors that are already resident on the device.
        points/rgb [N,3] f32, bidx [N] f32 frame index, offs [nb+1] host frame offsets; gt_labels: optional
        [N] uint8 device tensor used for the EE crop instead of the predicted labels (random-init weights give no
        usable EE). Returns (per-point labels uint8 on the device, pose dict on the host)."""
        nb = len(offs) - 1
        offs_dev = torch.as_tensor(np.asarray(offs), dtype=torch.int32).to(points.device)
        with _Stage(self, "segmentation"):
            rgbn = rgb if rgb_normalized else normalize_colors_(rgb, bidx, offs_dev)   # b200calib.ingest normalises
            labels, fld, out = segment_points(self.seg_model, points, rgbn, bidx, nb, self.cfg.seg_scale)
            del fld, out
        seg_labels = labels
        crop_src = labels if gt_labels is None else gt_labels
        with _Stage(self, "ee_cluster"):
            labels2, ee_idx, ee_offs = self.filter_ee(points, crop_src, offs_dev, nb)
        if gt_labels is None:
            seg_labels = labels2
        pose = self.pose_from_ee(points, rgbn, ee_idx, ee_offs, ee2base_poses, kp_conf_threshold)
        pose["ee_counts"] = (ee_offs[1:] - ee_offs[:-1]).numpy()
        pose["crop_labels"] = labels2   # device tensor: the labels the EE crop was taken from (2 = the kept EE cluster)
        return seg_labels, pose

    @torch.no_grad()
    def predict_batch(self, frames, ee2base_poses=None, gt_labels=None, kp_conf_threshold=None):
        """frames: list of (points [N,3] f32, rgb [N,3] f32) host arrays. gt_labels: optional list of label arrays
        used for the EE crop instead of the predicted labels (random-init weights give no usable EE)."""
        dev = torch.device("cuda")
        points, rgb, bidx, offs = batch_frames(frames, dev)
        gl = None
        if gt_labels is not None:
            gl = torch.as_tensor(np.concatenate(gt_labels).astype(np.uint8)).to(dev)
        seg_labels, pose = self.predict_device(points, rgb, bidx, offs, ee2base_poses, gl, kp_conf_threshold)
        results = self.assemble(seg_labels.cpu().numpy(), offs, pose, ee2base_poses)
        self.attach_key_points(results, frames, offs, pose)
        return results

    def predict_stream(self, batches, depth=2, fn=None, worker_init=None, **kw):
        """Throughput mode: iterate over batches (each a tuple (points, rgb, bidx, offs) of device tensors + host
        offsets, as predict_device takes them; or whatever `fn(item)` takes: fn runs in the worker, on its stream,
        instead of predict_device) with `depth` batches in flight, each on its own CUDA stream and host thread
        (`worker_init(w)` is called once in every worker thread). The small networks of batch i (rotation / key points on the EE crops: launch-bound, the host cannot
        keep the GPU busy) then run in the gaps of the large convolutions of batch i + 1 and vice versa, and every
        host sync of one batch is hidden behind the other's kernels. Yields (labels, pose) in batch order."""
        import queue
        import threading
        dev = torch.device("cuda", torch.cuda.current_device())
        main = torch.cuda.current_stream(dev)
        # the worker streams are kept for the life of the engine: PyTorch's caching allocator pools blocks per stream,
        # so fresh streams per call would cudaMalloc every activation buffer again on every call
        pool = self.__dict__.setdefault("_worker_streams", {}).setdefault(dev.index, [])
        while len(pool) < depth:
            pool.append(torch.cuda.Stream(dev))
        streams = pool[:depth]
        q_in, q_out = [queue.Queue() for _ in range(depth)], [queue.Queue() for _ in range(depth)]

        def worker(w):
            torch.cuda.set_device(dev)
            if worker_init is not None:
                worker_init(w)
            with torch.cuda.stream(streams[w]):
                while True:
                    item = q_in[w].get()
                    if item is None:
                        return
                    try:
                        streams[w].wait_stream(main)      # inputs produced on the caller's stream
                        out = fn(item) if fn is not None else self.predict_device(*item, **kw)
                        done = torch.cuda.Event()
                        done.record(streams[w])
                        q_out[w].put((out, done, None))
                    except Exception as e:                # noqa: BLE001 - handed to the consumer
                        q_out[w].put((None, None, e))

        threads = [threading.Thread(target=worker, args=(w,), daemon=True) for w in range(depth)]
        for t in threads:
            t.start()
        try:
            pending = 0
            it = iter(batches)
            nxt = 0
            head = 0
            exhausted = False
            while True:
                while not exhausted and pending < depth:
                    try:
                        item = next(it)
                    except StopIteration:
                        exhausted = True
                        break
                    q_in[nxt % depth].put(item)
                    nxt += 1
                    pending += 1
                if pending == 0:
                    break
                out, done, err = q_out[head % depth].get()
                head += 1
                pending -= 1
                if err is not None:
                    raise err
                main.wait_event(done)                     # consumers on the caller's stream see the results
                yield out
        finally:
            for w in range(depth):
                q_in[w].put(None)
            for t in threads:
                t.join(timeout=5)

    def attach_key_points(self, results, frames, offs, pose):
        """ResultDTO.key_points (app/inference_engine.py:327-345): (class, xyz) of every key point above the confidence
        threshold, from the host copy of the frames (predict_batch only)."""
        if pose.get("key_points") is None:
            return
        crop = pose["crop_labels"].cpu().numpy()
        probs, idx, _ = pose["key_points"]
        for j, f in enumerate(pose["ok_frames"]):
            r = results[f]
            if r.ee_pose is None:
                continue
            ee_pts = np.asarray(frames[f][0])[crop[offs[f]:offs[f + 1]] == 2]
            r.key_points = [(int(k), ee_pts[idx[j, k]]) for k in np.nonzero(probs[j] > pose["kp_threshold"])[0]]

    def host_sanity(self, results, frames, offs, pose):
        """cross-check of the device verdict: check_sanity of the host mirror (b200calib/sanity.py, pinned by the
        reference's own outputs) on the host copy of the frames -> list of (frame, verdict)."""
        from .sanity import check_sanity
        crop = pose["crop_labels"].cpu().numpy()
        out = []
        for j, f in enumerate(pose["ok_frames"]):
            r = results[f]
            if r.ee_pose is None:
                continue
            lab_f = crop[offs[f]:offs[f + 1]]
            out.append((int(f), bool(check_sanity(np.asarray(frames[f][0]), lab_f, pose["ee_pose_initial"][j],
                                                   r.key_points, self.cfg.sanity_min_ee_points,
                                                   self.cfg.sanity_kp_error_margin))))
        return out

    @staticmethod
    def assemble(seg_h, offs, pose, ee2base_poses=None):
        """host-side ResultDTO assembly (app/inference_engine.py:288-369). is_confident is the device verdict of
        check_sanity (False for every frame when PipelineConfig.sanity_check is off: unchecked frames never reach the
        calibration average)."""
        nb = len(offs) - 1
        results = [FrameResult(segmentation=seg_h[offs[i]:offs[i + 1]]) for i in range(nb)]
        conf = pose.get("confident")
        for j, f in enumerate(pose["ok_frames"]):
            r = results[f]
            if pose.get("ee_pose") is not None:
                r.ee_pose = pose["ee_pose"][j]
            if pose.get("kp_pose") is not None and pose["kp_valid"][j]:
                r.key_points_pose = pose["kp_pose"][j]
            if pose.get("icp_stats") is not None:
                r.icp_stats = pose["icp_stats"][j]
            if pose.get("vote_center") is not None:
                r.vote_center = pose["vote_center"][j]
            if ee2base_poses is not None and ee2base_poses[f] is not None:
                if r.ee_pose is not None:
                    r.base_pose = get_base2cam_pose(r.ee_pose, ee2base_poses[f])
                if r.key_points_pose is not None:
                    r.key_points_base_pose = get_base2cam_pose(r.key_points_pose, ee2base_poses[f])
            r.is_confident = bool(conf[j]) if (conf is not None and r.ee_pose is not None) else False
        return results


def _poses_to_matrices(poses):
    """[S,7] x,y,z,qw,qx,qy,qz (torch f64, device) -> [S,4,4] with the formula of utils/transformation.py:16-60."""
    q0, q1, q2, q3 = poses[:, 3], poses[:, 4], poses[:, 5], poses[:, 6]
    T = torch.zeros((poses.shape[0], 4, 4), dtype=torch.float64, device=poses.device)
    T[:, 0, 0] = 2 * (q0 * q0 + q1 * q1) - 1
    T[:, 0, 1] = 2 * (q1 * q2 - q0 * q3)
    T[:, 0, 2] = 2 * (q1 * q3 + q0 * q2)
    T[:, 1, 0] = 2 * (q1 * q2 + q0 * q3)
    T[:, 1, 1] = 2 * (q0 * q0 + q2 * q2) - 1
    T[:, 1, 2] = 2 * (q2 * q3 - q0 * q1)
    T[:, 2, 0] = 2 * (q1 * q3 - q0 * q2)
    T[:, 2, 1] = 2 * (q2 * q3 + q0 * q1)
    T[:, 2, 2] = 2 * (q0 * q0 + q3 * q3) - 1
    T[:, :3, 3] = poses[:, :3]
    T[:, 3, 3] = 1.0
    return T
