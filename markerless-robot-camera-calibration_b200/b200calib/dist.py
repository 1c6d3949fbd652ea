"""Frame-level data parallelism: frames are independent units, every rank holds a full replica
of the networks, and the only collective is one all-gather of the per-frame result records at the end.
One process per GPU (torchrun); backend nccl on GPUs, gloo in the CPU tests."""
import os

import numpy as np
import torch
import torch.distributed as dist

RECORD_WIDTH = 20  # frame id, ok flag, #ee points, ee_pose[7], kp_pose[7], icp fitness, icp rmse, icp iterations


def init_from_env(backend=None):
    """-> (rank, world, local_rank). No-op single-process defaults when WORLD_SIZE is unset or 1."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29511")
        backend = backend or ("nccl" if torch.cuda.is_available() else "gloo")
        if backend == "nccl":
            torch.cuda.set_device(local)
        dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, world, local


def shard_frames(n_frames, rank, world, weights=None):
    """frame ids owned by `rank`. Round-robin (f mod world), or greedy balancing by point count when weights
    (points per frame) are given: heaviest frame first onto the lightest rank, ties to the lowest rank."""
    if weights is None:
        return list(range(rank, n_frames, world))
    order = sorted(range(n_frames), key=lambda f: (-weights[f], f))
    load = [0] * world
    owner = {}
    for f in order:
        r = min(range(world), key=lambda i: (load[i], i))
        owner[f] = r
        load[r] += weights[f]
    return sorted(f for f, r in owner.items() if r == rank)


def pack_records(frame_ids, results):
    """FrameResult list -> [n, RECORD_WIDTH] float64 array."""
    rec = np.full((len(frame_ids), RECORD_WIDTH), np.nan)
    for i, (f, r) in enumerate(zip(frame_ids, results)):
        rec[i, 0] = f
        rec[i, 1] = 1.0 if r.ee_pose is not None else 0.0
        rec[i, 2] = float((r.segmentation == 2).sum())
        if r.ee_pose is not None:
            rec[i, 3:10] = r.ee_pose
        if r.key_points_pose is not None:
            rec[i, 10:17] = r.key_points_pose
        if r.icp_stats is not None:
            rec[i, 17:20] = r.icp_stats[:3]
    return rec


def gather_records(local_records, device=None):
    """all-gather variable-length [n_i, RECORD_WIDTH] record blocks; returns them sorted by frame id."""
    t = torch.as_tensor(local_records, dtype=torch.float64)
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        out = t
    else:
        world = dist.get_world_size()
        if device is None:
            device = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else "cpu"
        n = torch.tensor([t.shape[0]], dtype=torch.int64, device=device)
        sizes = [torch.zeros_like(n) for _ in range(world)]
        dist.all_gather(sizes, n)
        nmax = int(max(s.item() for s in sizes))
        pad = torch.full((nmax, RECORD_WIDTH), float("nan"), dtype=torch.float64, device=device)
        pad[: t.shape[0]] = t.to(device)
        gathered = torch.empty((world * nmax, RECORD_WIDTH), dtype=torch.float64, device=device)
        dist.all_gather_into_tensor(gathered, pad)
        parts = [gathered[r * nmax: r * nmax + int(sizes[r].item())] for r in range(world)]
        out = torch.cat(parts).cpu()
    order = torch.argsort(out[:, 0])
    return out[order].numpy()
