"""Per-point heads with the names of utils/output.py, batched over frames on the GPU (K5, K7, K8).

Segments: frames (or EE crops) are stored back to back; `seg_offsets` [S+1] int32 gives each one's rows."""
import numpy as np
import torch

from MinkowskiEngine._lib import lib, check, ptr, stream, B2MEError
from MinkowskiEngine.core import _count


def _offsets(seg_offsets, device):
    return torch.as_tensor(seg_offsets, dtype=torch.int32, device=device).contiguous()


def get_segmentations_from_tensor_field(field):
    """utils/output.py:67-73: (preds, conf) = arg-max class and sigmoid(max logit) per point."""
    logits = field.features.float()
    conf, preds = logits.max(1)
    return preds.cpu().numpy(), torch.sigmoid(conf).cpu().numpy()


def largest_cluster_mask(points, seg_offsets, dist=0.06):
    """K7 batched: points [n,3] f32 CUDA, seg_offsets [S+1]. Returns (mask [n] uint8, sizes [S] int32):
    mask = 1 on the largest single-linkage cluster (merge distance < dist) of every segment."""
    points = points.to(torch.float32).contiguous()
    dev = points.device
    n = points.shape[0]
    offs = _offsets(seg_offsets, dev)
    S = offs.numel() - 1
    mask = torch.zeros((max(n, 1),), dtype=torch.uint8, device=dev)
    sizes = torch.zeros((max(S, 1),), dtype=torch.int32, device=dev)
    for s0 in range(0, S, 1024):  # the kernel's key packs the segment id in 10 bits
        s1 = min(S, s0 + 1024)
        r0, r1 = int(seg_offsets[s0]), int(seg_offsets[s1])
        sub = (offs[s0:s1 + 1] - r0).contiguous()
        ws = torch.empty((lib.b2me_cluster_workspace_bytes(r1 - r0, s1 - s0),), dtype=torch.uint8, device=dev)
        check(lib.b2me_largest_cluster(ptr(points[r0:r1]), ptr(sub), s1 - s0, r1 - r0, float(dist), ptr(mask[r0:]),
                                       ptr(sizes[s0:]), ptr(ws), ws.numel(), stream()), "largest_cluster")
        _count(18)  # cells (1) + unique (7) + count, scan x3, fill (5) + link, flatten, root, best, mask (5)
    return mask[:n], sizes[:S]


class ClusterUtil:
    """utils/output.py:13-28 (sklearn AgglomerativeClustering, single linkage, distance_threshold) on K7."""

    def __init__(self, dist=0.06, linkage="single"):
        if linkage != "single":
            raise NotImplementedError("only single linkage is implemented")
        self.dist = dist

    def get_largest_cluster(self, points):
        pts = torch.as_tensor(np.asarray(points, dtype=np.float32)) if not torch.is_tensor(points) else points
        pts = pts.to("cuda", torch.float32)
        mask, _ = largest_cluster_mask(pts, [0, pts.shape[0]], self.dist)
        return torch.nonzero(mask).flatten().cpu().numpy()


def key_point_predictions_batched(logits, seg_offsets):
    """K8a: per (segment, class) the best soft-max probability over the segment's points and the lowest row
    index reaching it. logits [n,K] f32 CUDA. Returns (best_prob [S,K] f32, best_idx [S,K] i32 global rows)."""
    logits = logits.to(torch.float32).contiguous()
    dev = logits.device
    offs = _offsets(seg_offsets, dev)
    S = offs.numel() - 1
    K = logits.shape[1]
    bp = torch.empty((S, K), dtype=torch.float32, device=dev)
    bi = torch.empty((S, K), dtype=torch.int32, device=dev)
    check(lib.b2me_keypoint_reduce(ptr(logits), K, ptr(offs), S, ptr(bp), ptr(bi), stream()), "keypoint_reduce")
    _count(1)
    return bp, bi


def get_key_point_predictions(logits, conf_th=0.999):
    """utils/output.py:81-87 for one cloud: (idx, classes, probs) of the classes whose best point beats conf_th."""
    logits = logits.to("cuda")
    bp, bi = key_point_predictions_batched(logits, [0, logits.shape[0]])
    bp, bi = bp[0].cpu(), bi[0].cpu().numpy()
    classes = np.where(bp > conf_th)[0]
    return bi[classes], classes, bp[classes]


def vote_centers_batched(logits, points, seg_offsets, col=1, topk=8):
    """K8b: mean coordinate of the topk rows with the largest logits[:, col] per segment -> [S,3] f32."""
    logits = logits.to(torch.float32).contiguous()
    points = points.to(torch.float32).contiguous()
    dev = logits.device
    offs = _offsets(seg_offsets, dev)
    S = offs.numel() - 1
    out = torch.empty((S, 3), dtype=torch.float32, device=dev)
    check(lib.b2me_vote_center(ptr(logits), logits.shape[1], col, ptr(points), ptr(offs), S, topk, ptr(out), stream()),
          "vote_center")
    _count(1)
    return out


def get_pred_center(out, coords, ee_r=0.03, q=None):
    """utils/output.py:45-64 for one cloud; `out` [n,C] logits, `coords` [n,3]; q = W,X,Y,Z."""
    from .transformation import get_quaternion_rotation_matrix
    out = torch.as_tensor(out).to("cuda")
    pts = torch.as_tensor(np.asarray(coords, dtype=np.float32)).to("cuda")
    center = vote_centers_batched(out, pts, [0, out.shape[0]])[0].cpu().numpy()
    if q is not None:
        qn = np.asarray(q.detach().cpu() if torch.is_tensor(q) else q, dtype=np.float32).reshape(-1)
        qn = qn / np.linalg.norm(qn)  # the reference's torch helper divides by |q|^2 inside the matrix formula
        center = center + (get_quaternion_rotation_matrix(qn, switch_w=False) @ np.array([-ee_r, 0, 0],
                                                                                         dtype=np.float32))
    return center


def translation_magic_batched(points, seg_offsets, quats_wxyz, x_offset=-0.015):
    """predict_translation with magic_enabled (app/inference_engine.py:459-489) for S segments -> [S,3] f64."""
    points = points.to(torch.float32).contiguous()
    dev = points.device
    offs = _offsets(seg_offsets, dev)
    S = offs.numel() - 1
    q = quats_wxyz.to(dev, torch.float32).contiguous()
    out = torch.empty((S, 3), dtype=torch.float64, device=dev)
    check(lib.b2me_translation_magic(ptr(points), ptr(offs), S, ptr(q), float(x_offset), ptr(out), stream()),
          "translation_magic")
    _count(1)
    return out


def sanity_check_batched(points, seg_offsets, ee_poses, kp_prob=None, kp_xyz=None, kp_threshold=0.75,
                         min_num_of_ee_points=2048, kp_error_margin=0.05):
    """InferenceEngine.check_sanity (app/inference_engine.py:246-279) for S EE crops on the device -> [S] uint8.
    points [n,3] f32 (the crops, concatenated), seg_offsets [S+1], ee_poses [S,7] f64 x,y,z,qw,qx,qy,qz (before the ICP
    refinement), kp_prob [S,K] f32 / kp_xyz [S,K,3] f32: the per-class best key-point probabilities and coordinates
    (a class counts when its probability exceeds kp_threshold)."""
    points = points.to(torch.float32).contiguous()
    dev = points.device
    offs = _offsets(seg_offsets, dev)
    S = offs.numel() - 1
    poses = ee_poses.to(dev, torch.float64).contiguous()
    K = 0
    if kp_prob is not None:
        kp_prob = kp_prob.to(dev, torch.float32).contiguous()
        kp_xyz = kp_xyz.to(dev, torch.float32).contiguous()
        K = kp_prob.shape[1]
    out = torch.zeros((max(S, 1),), dtype=torch.uint8, device=dev)
    check(lib.b2me_sanity_check(ptr(points), ptr(offs), S, ptr(poses), ptr(kp_prob), ptr(kp_xyz), K, float(kp_threshold),
                                int(min_num_of_ee_points), float(kp_error_margin), ptr(out), stream()), "sanity_check")
    _count(1)
    return out[:S]


def normalize_colors_batched(rgb, bidx, frame_offsets):
    """utils/preprocess.py:20-37 per frame of a batch on the device (no host round trip): rgb [n,3] f32, bidx [n] f32
    frame index, frame_offsets [F+1] -> [n,3] f32."""
    rgb = rgb.to(torch.float32).contiguous()
    dev = rgb.device
    n = rgb.shape[0]
    offs = _offsets(frame_offsets, dev)
    F = offs.numel() - 1
    out = torch.empty_like(rgb)
    ws = torch.empty((max(F, 1) * 24,), dtype=torch.uint8, device=dev)
    check(lib.b2me_normalize_colors(ptr(rgb), ptr(bidx.contiguous()), n, ptr(offs), F, ptr(out), ptr(ws), ws.numel(),
                                    stream()), "normalize_colors")
    _count(3)
    return out
