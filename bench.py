#!/usr/bin/env python
"""bench.py — segmented-and-posed frames/s of the markerless-calibration inference hot path on B200.

One "step" = one pass of the whole per-batch pipeline over `--frames` synthetic 640x480 Kinect-shaped frames per
GPU (BASELINE.json configs[1] + [2]): voxelise (K1) -> kernel maps (K2/K3) -> MinkUNet18D segmentation forward in
bf16 on tcgen05 (K4) with the 256 -> 1024 -> 3 head fused into the last GEMM's epilogue (K5) -> per-point labels ->
largest EE cluster (K7) -> RobotNetEncode rotation (K1-K4, K6) -> "magic" translation -> key-point network +
per-class reduction (K8) -> batched Kabsch (K9) -> sanity check -> ICP x2 (K10, one persistent cluster launch).
Random-init weights of the named architectures, synthetic data (dataset and checkpoints are unpublished).

  python bench.py --gpus N --steps K --warmup W            # B200 arm (under torchrun for N > 1)
  python bench.py --impl reference --gpus N ...            # CPU oracle of the reference path on the host cores
  python bench.py --config icp1k                           # BASELINE.json configs[3]: 1 000-frame EE ICP refinement
  python bench.py --config sweep [--gpus N]                # BASELINE.json configs[4]: voxel-size sweep + wide scene
  python bench.py --strong-frames 256 --gpus N             # strong scaling: fixed uneven workload, greedy sharding

Prints ONE JSON line (rank 0). `value` = frames/s with inputs resident in HBM; `e2e` = the same through the public
API with pinned host buffers (H2D of the inputs, D2H of labels and poses inside the timed region; the H2D copy of the
next step is issued on a copy stream under the current step's kernels). The `roofline` object is measured in the same
single-stream pass as `value` (CUDA events around every convolution launch); with `--depth` > 1 batches in flight
(BatchedInferenceEngine.predict_stream) `value` comes from a separate pass, because events around a launch only measure
that kernel when no other stream competes for the GPU.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "markerless-robot-camera-calibration_b200")
for _p in (ROOT, PKG):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import numpy as np  # noqa: E402

SEED = 13  # TEST.seed of config/default.yaml:112


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=6)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="pipeline", choices=["pipeline", "icp1k", "sweep"],
                    help="pipeline = BASELINE.json configs[1]+[2] (the headline); icp1k = configs[3]; sweep = configs[4]")
    ap.add_argument("--frames", type=int, default=32, help="frames per GPU per step (batch)")
    ap.add_argument("--width", type=int, default=640)
    ap.add_argument("--height", type=int, default=480)
    ap.add_argument("--scale", type=float, default=200.0, help="voxels per metre (200 = 5 mm, production)")
    ap.add_argument("--depth", type=int, default=1,
                    help="batches in flight of the timed loops (BatchedInferenceEngine.predict_stream: one CUDA stream + "
                         "host thread per batch). Default 1: the convolutions are persistent whole-GPU kernels, so a second "
                         "stream only queues its small kernels behind them (measured: depth 2 = 98 vs 136 frames/s)")
    ap.add_argument("--crop", default="both", choices=["gt", "pred", "both"],
                    help="EE crop of the timed steps: ground-truth labels (stable workload), predicted labels, or both "
                         "(the predicted-crop pass is reported as `pred_crop`)")
    ap.add_argument("--strong-frames", type=int, default=0,
                    help="strong scaling: this many frames IN TOTAL with uneven point counts, sharded over the ranks by "
                         "greedy balancing; gather + calibrate inside the timed region")
    ap.add_argument("--vote", action="store_true",
                    help="also run the vote head (RobotNetVote + get_pred_center, model/robotnet_vote.py, utils/output.py:"
                         "45-64) on the EE crops inside every step; InferenceEngine.predict does not call it, so it is off "
                         "in the headline workload")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-icp", action="store_true")
    ap.add_argument("--stages", action="store_true", help="one extra (untimed) step with synchronised per-stage times")
    ap.add_argument("--dtype", default="bf16", choices=["bf16", "tf32", "f32"])
    ap.add_argument("--tc-path", default="tma", choices=["cpasync", "tma"], help="operand path of k_spconv_tc")
    ap.add_argument("--no-rot128", action="store_true", help="A/B: single 384-column accumulator (round-1 layout)")
    ap.add_argument("--no-fuse-head", action="store_true", help="A/B: two-launch head (hidden activation in HBM)")
    ap.add_argument("--kp-backbone", default="minkunet", choices=["minkunet", "pointnet2"],
                    help="key-point network: 6-class MinkUNet18D (default; both arms) or the reference's default "
                         "PointNet2SSG branch on GPU-native FPS / ball query / 3-NN (B200 arm only)")
    ap.add_argument("--mask-block", type=int, default=None, help="rows per locality block of the K3b mask sort (0 = global)")
    ap.add_argument("--mask-morton", action="store_true", help="A/B: Morton order inside a mask group of the k3 maps")
    ap.add_argument("--mask-two-level", action="store_true", help="A/B: two-level K3b mask-sort keys (default: one-level)")
    ap.add_argument("--cprofile", default=None, help="one extra (untimed) step under cProfile: host-side launch cost (text)")
    ap.add_argument("--torch-profile", default=None,
                    help="one extra (untimed) step under torch.profiler: per-kernel device times + GPU busy fraction (text)")
    ap.add_argument("--profile-range", action="store_true",
                    help="one extra device-resident step between cudaProfilerStart / Stop at the end (ncu "
                         "--profile-from-start off then captures exactly one warm step)")
    ap.add_argument("--conv-table", default=None, help="write the per-convolution census + CUDA-event times here (JSON)")
    ap.add_argument("--report", default=None, help="write the evaluation record of the last step here (JSON)")
    return ap.parse_args()


def _gen_frame(a):
    from b200calib.synthetic import make_frame
    seed, w, h = a[:3]
    f = make_frame(seed, width=w, height=h, wide=bool(a[3]) if len(a) > 3 else False)
    rgb = (np.round(f["rgb"] * 255.0) / 255.0).astype(np.float32)   # 8-bit colours, as a camera delivers them
    return f["points"], rgb, f["labels"].astype(np.uint8), f["ee_pose_wxyz"]


def make_workload(n_frames, rank, width, height, wide=False, seeds=None, sizes=None):
    """n_frames distinct seeded frames for this rank (generated before CUDA is touched, in worker processes)."""
    import multiprocessing as mp
    if seeds is None:
        seeds = [SEED * 1000 + rank * 4096 + i for i in range(n_frames)]
    if sizes is None:
        sizes = [(width, height)] * len(seeds)
    jobs = [(s, w, h, wide) for s, (w, h) in zip(seeds, sizes)]
    nproc = max(1, min(8, (os.cpu_count() or 8) // max(1, int(os.environ.get("WORLD_SIZE", "1")))))
    if nproc > 1 and len(jobs) > 1:
        with mp.get_context("fork").Pool(nproc) as pool:
            return pool.map(_gen_frame, jobs)
    return [_gen_frame(j) for j in jobs]


def load_ncu_traffic(frames):
    """dram__bytes_read.sum + dram__bytes_write.sum per k_spconv_tc launch (mean over the launches of one step) from a
    COMMITTED ncu capture of this command (profiles/*_k_spconv_tc_dram_traffic.json, newest round first); None for
    another batch size or when no capture is committed. It is ncu evidence, not something this run measured."""
    for name in ("r02_k_spconv_tc_dram_traffic.json", "r01_k_spconv_tc_dram_traffic.json"):
        try:
            with open(os.path.join(ROOT, "profiles", name)) as fp:
                t = json.load(fp)
            if int(t.get("frames", -1)) == int(frames):
                return float(t["dram_bytes_per_launch"]), name
        except (OSError, ValueError, KeyError):
            continue
    return None, None


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], tf_burst=d["bf16_tflops"], tf_sustained=d.get("bf16_tflops_sustained",
                                                                                 d["bf16_tflops"]), src="measured")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sustained=1400.0, src="fallback")


def measure_tf32_peak(dev, seconds=1.5):
    """cuBLAS tf32 GEMM 8192^3 on this box: (best of 10, back-to-back rate over `seconds`): the roofline denominator of
    the tf32 tensor-core mode (MEASURED_PEAKS.json holds bf16 only)."""
    import torch
    old = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = True
    try:
        n = 8192
        a = torch.randn(n, n, device=dev)
        b = torch.randn(n, n, device=dev)
        for _ in range(3):
            a @ b
        torch.cuda.synchronize()
        best = 1e9
        for _ in range(10):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            a @ b
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = max(10, int(seconds * 1e3 / best))
        e0.record()
        for _ in range(reps):
            a @ b
        e1.record()
        torch.cuda.synchronize()
        return 2.0 * n ** 3 / (best * 1e-3) / 1e12, 2.0 * n ** 3 * reps / (e0.elapsed_time(e1) * 1e-3) / 1e12
    finally:
        torch.backends.cuda.matmul.allow_tf32 = old


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if self.proc is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        time.sleep(0.25)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx = float(f[1])
            except ValueError:
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return dict(sm_mhz=float(np.median(sm)) if sm else None, sm_max_mhz=mx, reasons=sorted(reasons),
                    samples=len(sm))


def build_models(ME, kp_classes=6, kp_backbone="minkunet"):
    import torch
    from b200calib.models import make_models, randomize_bn_stats
    torch.manual_seed(SEED)
    M = make_models(ME)
    seg = randomize_bn_stats(M.RobotNetSegmentation(3, num_classes=3), SEED).eval()
    rot = randomize_bn_stats(M.RobotNetEncode(3, 7), SEED + 1).eval()
    if kp_backbone == "pointnet2":
        from b200calib.pointnet2 import PointNet2SSG
        kp = PointNet2SSG(num_classes=kp_classes, in_channels=6).eval()
    else:
        kp = randomize_bn_stats(M.RobotNetSegmentation(3, num_classes=kp_classes), SEED + 2).eval()
    return seg, rot, kp


def load_cad():
    """the reference's ICP source: the xyz of app/hand_files/hand.pcd (4480 points; committed fixture)."""
    return np.load(os.path.join(ROOT, "tests", "golden", "cad_hand_points.npz"))["xyz"].astype(np.float32)


# ---------------------------------------------------------------------------------------------- reference arm
def run_reference(args, rank, world):
    """the reference's own per-frame CPU path, restated (oracle/pipeline.py; the real one needs MinkowskiEngine 0.5.4
    + Open3D, absent here), on all host threads. Each step = ONE frame of the workload (a bounded sample); warm-up and
    step counts are the ones asked for, like the B200 arm."""
    if rank != 0:
        return
    import torch
    import oracle.MinkowskiEngine as OME
    from oracle import pipeline as op
    from b200calib.synthetic import ee_surface_cloud
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    if args.config == "icp1k":
        return run_icp1k_reference(args, cores)
    seg, rot, kp = build_models(OME)
    frames = make_workload(max(1, min(args.frames, args.warmup + args.steps)), 0, args.width, args.height)
    cad = ee_surface_cloud(4096, SEED)
    cfg = dict(seg_scale=args.scale, icp_enabled=not args.no_icp)
    models = dict(seg=seg, rot=rot, kp=kp)

    def one(i):
        p, c, l = frames[i % len(frames)][:3]
        return op.predict_frame(models, cad, p, c, cfg, gt_labels=l)

    for i in range(args.warmup):
        one(i)
    t0 = time.perf_counter()
    for i in range(args.steps):
        one(args.warmup + i)
    dt = time.perf_counter() - t0
    fps = args.steps / dt
    sample = (f"{args.steps} steps of ONE {args.width}x{args.height} frame each (batch 1 fp32, the reference's mode) after "
              f"{args.warmup} warm-up frames")
    print(json.dumps({
        "impl": "reference", "metric": "segmented-and-posed frames/s", "value": fps, "unit": "frames/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args),
        "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def workload_config(args):
    kpn = "6-key-point MinkUNet18D" if args.kp_backbone == "minkunet" else "6-key-point PointNet2SSG on 2048 points"
    return {"workload": (f"segmentation forward (MinkUNet18D, 3 classes) + EE pose (RobotNetEncode rotation, magic "
                         f"translation, {kpn} + Kabsch, sanity check, ICP x2) on {args.frames} synthetic "
                         f"{args.width}x{args.height} Kinect-shaped frames per GPU, voxel {1.0 / args.scale * 1000:.1f} mm "
                         f"(seg/rot), 1.25 mm (key points); BASELINE.json configs[1]+[2]"),
            "frames_per_gpu": args.frames, "points_per_frame": "~3.0e5", "voxel_m": 1.0 / args.scale,
            "ee_crop": "ground-truth labels (random-init weights give no usable EE prediction); `pred_crop` = the same "
                       "steps with the crop taken from the predicted labels",
            "batches_in_flight": args.depth,
            "l2": "inputs and activations exceed L2 (126 MB) every step; no explicit flush",
            "parallelism": "frames sharded across GPUs, one final all_gather of pose records"}


# ---------------------------------------------------------------------------------------------- B200 arm
class Workbench:
    """models + engine + this rank's frames in pinned host memory and on the device."""

    def __init__(self, args, rank, world, local, frames):
        import torch
        import MinkowskiEngine as ME
        from b200calib.pipeline import BatchedInferenceEngine, PipelineConfig
        from b200calib.ingest import pack_xyzrgb
        self.torch, self.ME, self.args = torch, ME, args
        self.dev = torch.device("cuda", local)
        ME.set_compute_dtype({"bf16": torch.bfloat16, "tf32": "tf32", "f32": torch.float32}[args.dtype])
        ME.set_tc_operand_path(args.tc_path)
        ME.set_tc_rot128(not args.no_rot128)
        ME.set_fuse_head(not args.no_fuse_head)
        if args.mask_block is not None:
            ME.set_mask_sort_block(args.mask_block)
        if args.mask_morton:
            ME.set_mask_sort_morton(True)
        if args.mask_two_level:
            ME.set_mask_sort_two_level(True)
        self.seg, self.rot, self.kp = [m.to(self.dev) for m in build_models(ME, kp_backbone=args.kp_backbone)]
        # ICP source of the pipeline configs: a cloud on the surface of the rendered EE box (both arms use it; the
        # reference's hand.pcd belongs to the real gripper and is the source of --config icp1k)
        from b200calib.synthetic import ee_surface_cloud
        self.cad = torch.from_numpy(ee_surface_cloud(4096, SEED)).to(self.dev)
        self.cfg = PipelineConfig(seg_scale=args.scale, icp_enabled=not args.no_icp)
        vote = None
        if getattr(args, "vote", False):
            from b200calib.models import make_models, randomize_bn_stats
            torch.manual_seed(SEED + 3)
            vote = randomize_bn_stats(make_models(ME).RobotNetVote(3, num_classes=2), SEED + 3).eval().to(self.dev)
        self.eng = BatchedInferenceEngine(self.seg, self.rot, self.kp, cad_points=self.cad, config=self.cfg,
                                          vote_model=vote)
        self.frames = frames
        self.pack_xyzrgb = pack_xyzrgb

    def stage(self, frames):
        """one batch: device-resident tensors + pinned host records (the "host buffers" of the e2e number)."""
        torch, dev = self.torch, self.dev
        counts = [len(f[0]) for f in frames]
        offs = np.zeros(len(frames) + 1, dtype=np.int64)
        np.cumsum(counts, out=offs[1:])
        N = int(offs[-1])
        b = dict(offs=offs, N=N, n=len(frames))
        pts = np.concatenate([f[0] for f in frames])
        rgb = np.concatenate([f[1] for f in frames])
        lab = np.concatenate([f[2] for f in frames])
        b["d_pts"] = torch.from_numpy(pts).to(dev)
        b["d_rgb"] = torch.from_numpy(rgb).to(dev)
        b["d_bidx"] = torch.from_numpy(np.repeat(np.arange(len(frames), dtype=np.float32), counts)).to(dev)
        b["d_lab"] = torch.from_numpy(lab).to(dev)
        # e2e host buffers: PointCloud2-style records (x, y, z, PCL-packed rgb; 16 B per point), the wire format the
        # reference's live source delivers (app/freenect_data_engine.py:74-81); b200calib.ingest unpacks / normalises /
        # ROI-filters them on the device
        b["h_rec"] = torch.from_numpy(np.concatenate([self.pack_xyzrgb(f[0], np.round(f[1] * 255.0))
                                                      for f in frames])).pin_memory()
        b["h_lab"] = torch.from_numpy(lab).pin_memory()
        b["offs32"] = offs.astype(np.int32)
        return b


def run_b200(args, rank, world, local):
    if args.config == "icp1k":
        return run_icp1k(args, rank, world, local)
    if args.config == "sweep":
        return run_sweep(args, rank, world, local)
    if args.strong_frames:
        return run_strong(args, rank, world, local)
    frames = make_workload(args.frames, rank, args.width, args.height)   # before CUDA init (fork pool)
    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py (B200 arm) needs a CUDA device; there is no CPU fallback")
    torch.cuda.set_device(local)
    import MinkowskiEngine as ME
    from b200calib import dist as bdist
    from b200calib.ingest import ingest_clouds
    bdist.init_from_env("nccl" if world > 1 else None)
    wb = Workbench(args, rank, world, local, frames)
    eng, dev = wb.eng, wb.dev
    B = wb.stage(frames)
    torch.cuda.synchronize()
    depth = max(1, args.depth)
    h_seg = [torch.empty((B["N"],), dtype=torch.uint8).pin_memory() for _ in range(depth + 1)]

    def step_device(gt=True):
        return eng.predict_device(B["d_pts"], B["d_rgb"], B["d_bidx"], B["offs"], gt_labels=B["d_lab"] if gt else None)

    copy_stream = torch.cuda.Stream(dev)

    def stage_in():
        """H2D of one step's inputs from pinned host memory on the copy stream (runs under the previous step's kernels)"""
        with torch.cuda.stream(copy_stream):
            r = B["h_rec"].to(dev, non_blocking=True)
            g = B["h_lab"].to(dev, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(copy_stream)
        return r, g, ev

    def step_e2e(slot, staged=None):
        r, g, ev = staged if staged is not None else stage_in()
        cur = torch.cuda.current_stream()
        cur.wait_event(ev)
        r.record_stream(cur)
        g.record_stream(cur)
        p, c, b, o = ingest_clouds(r, B["offs32"])     # synthetic frames have no invalid pixels: nothing is dropped
        labels, pose = eng.predict_device(p, c, b, o, gt_labels=g, rgb_normalized=True)
        h_seg[slot].copy_(labels, non_blocking=True)
        cur.synchronize()
        return eng.assemble(h_seg[slot].numpy(), o, pose)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn_items, fn, worker_init=None):
        """K steps with `depth` batches in flight; device time from an event on the caller's stream before the first
        batch is issued to one after the last result has been joined (predict_stream orders both)."""
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record()
        outs = list(eng.predict_stream(fn_items, depth=depth, fn=fn, worker_init=worker_init))
        e1.record()
        barrier()
        return e0.elapsed_time(e1), outs

    # ---- warm-up (also: census of the algorithmic work of every convolution launch, once)
    recs_census = ME.set_profile("census")
    step_device()
    ME.set_profile(None)
    for _ in range(max(args.warmup - 1, 0)):
        step_device()
    barrier()

    # ---- timed, device-resident: K steps on ONE stream with CUDA events around every convolution launch (the roofline
    #      numbers come from the same pass as `value` when depth == 1)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    recs_ev = ME.set_profile("events")
    ME.reset_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        labels, pose = step_device()
    e1.record()
    barrier()
    ms_serial = e0.elapsed_time(e1)
    launches = ME.launch_count()
    ME.set_profile(None)
    ms_dev = ms_serial
    if depth > 1:   # `depth` batches in flight: a separate pass (events around a launch would time the other stream too)
        ms_dev, outs = timed(range(args.steps), lambda i: step_device())
        labels, pose = outs[-1]
    clocks = sampler.stop() if rank == 0 else None

    # ---- timed: end to end through the public API with pinned host buffers; the H2D copy of step i + 1 is issued on a
    #      copy stream before step i's kernels (copy engines run under the kernels), every copy is inside the timed region
    step_e2e(depth)
    if depth > 1:
        ms_e2e, results_all = timed(range(args.steps), lambda i: step_e2e(i % depth))
        results = results_all[-1]
    else:
        e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e2.record()
        nxt = stage_in()
        for i in range(args.steps):
            staged, nxt = nxt, (stage_in() if i + 1 < args.steps else None)
            results = step_e2e(0, staged)
        e3.record()
        barrier()
        ms_e2e = e2.elapsed_time(e3)
    posed = sum(1 for r in results if r.ee_pose is not None)
    confident = sum(1 for r in results if r.is_confident)

    # ---- the same steps with the EE crop taken from the PREDICTED labels ("report both")
    pred_crop = None
    if args.crop in ("pred", "both"):
        step_device(gt=False)
        e4, e5 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e4.record()
        for _ in range(args.steps):
            _, pp = step_device(gt=False)
        e5.record()
        barrier()
        ms_pred = e4.elapsed_time(e5)
        pred_crop = {"value": args.frames * world * args.steps / (ms_pred * 1e-3), "unit": "frames/s",
                     "ms_per_step": ms_pred / args.steps, "frames_posed_per_step": int(len(pp["ok_frames"])),
                     "ee_points_per_frame_median": float(np.median(pp["ee_counts"])),
                     "note": "random-init weights: the predicted EE set is whatever class 2 wins, not an end effector"}

    stage_ms = None
    if args.stages:
        eng.stage_times = {}
        step_device()
        stage_ms = {k: round(v, 3) for k, v in eng.stage_times.items()}
        eng.stage_times = None

    if args.profile_range and rank == 0:
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
        step_device()
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()

    if args.cprofile and rank == 0:
        # developer aid: where the HOST spends its time while launching one step (the GPU idles when it falls behind)
        import cProfile
        import pstats
        torch.cuda.synchronize()
        pr = cProfile.Profile()
        t0 = time.perf_counter()
        pr.enable()
        step_device()
        pr.disable()
        t_launch = (time.perf_counter() - t0) * 1e3
        torch.cuda.synchronize()
        t_total = (time.perf_counter() - t0) * 1e3
        with open(args.cprofile, "w") as fp:
            fp.write(f"one step of {args.frames} frames: host returned after {t_launch:.1f} ms, GPU done after {t_total:.1f} ms\n")
            pstats.Stats(pr, stream=fp).sort_stats("cumulative").print_stats(45)
            pstats.Stats(pr, stream=fp).sort_stats("tottime").print_stats(30)

    if args.torch_profile and rank == 0:
        # developer aid: device time per kernel of ONE step as CUPTI sees it (no replay, warm caches) and how much of
        # the step the GPU had a kernel running (the rest is host launch / sync latency)
        from torch.profiler import profile, ProfilerActivity
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            step_device()
            torch.cuda.synchronize()
        wall_ms = (time.perf_counter() - t0) * 1e3
        ka = prof.key_averages()
        busy_ms = sum(getattr(k, "device_time_total", getattr(k, "cuda_time_total", 0)) for k in ka) / 1e3
        with open(args.torch_profile, "w") as fp:
            fp.write(f"one step of {args.frames} frames under torch.profiler: wall {wall_ms:.1f} ms, "
                     f"sum of kernel device time {busy_ms:.1f} ms ({100 * busy_ms / wall_ms:.1f} % busy)\n")
            fp.write(ka.table(sort_by="cuda_time_total", row_limit=40, max_name_column_width=60))

    # ---- the only collective: all-gather of the per-frame records (outside the per-step loop, as in production)
    recs = bdist.pack_records(list(range(rank * args.frames, (rank + 1) * args.frames)), results)
    allrec = bdist.gather_records(recs)
    t = torch.tensor([ms_dev, ms_e2e, ms_serial], dtype=torch.float64, device=dev)
    if world > 1:
        tl = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(tl, t)
        per_rank = torch.stack(tl).cpu().numpy()
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    else:
        per_rank = t.cpu().numpy()[None]
    ms_dev, ms_e2e, ms_serial = t.tolist()

    # ---- roofline of the dominant kernel (k_spconv_tc): FLOPs from the census / CUDA-event durations (serial pass)
    per_step = len(recs_census)
    tc_flops = tc_ms = simt_ms = tc_exec = 0.0
    n_tc = 0
    torch.cuda.synchronize()
    for i, ev in enumerate(recs_ev):
        c = recs_census[i % per_step]
        ms = ev[0].elapsed_time(ev[1])
        if c["kind"] == "tc":
            tc_flops += 2.0 * c["pairs"] * c["Cin"] * c["Cout"]
            tc_exec += 2.0 * c.get("passes", 0) * 256.0 * c["Cin"] * c["Cout"]
            tc_ms += ms
            n_tc += 1
        else:
            simt_ms += ms
    if args.conv_table and rank == 0:
        tab = []
        for i in range(per_step):
            c = dict(recs_census[i])
            ms = [recs_ev[j][0].elapsed_time(recs_ev[j][1]) for j in range(i, len(recs_ev), per_step)]
            c["ms"] = float(np.mean(ms))
            # time between the end of the previous convolution launch and the start of this one (same stream): the
            # coordinate kernels at a level change, the other stages between two networks, or the host falling behind
            gaps = [recs_ev[j - 1][1].elapsed_time(recs_ev[j][0]) for j in range(i, len(recs_ev), per_step) if j % per_step]
            c["gap_before_ms"] = float(np.mean(gaps)) if gaps else 0.0
            c["tflops"] = 2.0 * c["pairs"] * c["Cin"] * c["Cout"] / (c["ms"] * 1e-3) / 1e12
            c["fill"] = c["pairs"] / max(1, c["K"] * c["V_out"])
            if c.get("passes"):
                c["row_efficiency"] = c["pairs"] / (c["passes"] * 256.0)
                c["mma_tflops"] = 2.0 * c["passes"] * 256.0 * c["Cin"] * c["Cout"] / (c["ms"] * 1e-3) / 1e12
            tab.append(c)
        json.dump(tab, open(args.conv_table, "w"))
    peaks = load_peaks()
    roof = None
    if tc_ms > 0:
        ach = tc_flops / (tc_ms * 1e-3) / 1e12
        peak, peak_src = peaks["tf_sustained"], f"{peaks['src']} bf16 sustained (burst {peaks['tf_burst']})"
        if args.dtype == "tf32" and rank == 0:
            burst, sus = measure_tf32_peak(dev)
            peak, peak_src = sus, f"cuBLAS tf32 8192^3 measured in this run: sustained {sus:.1f} (burst {burst:.1f})"
        traffic, traffic_src = load_ncu_traffic(args.frames)
        roof = {"kernel": "k_spconv_tc (tcgen05 gather-GEMM sparse convolution"
                          + (", kind::tf32)" if args.dtype == "tf32" else ", kind::f16 bf16)"),
                "bound": "tensor", "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak,
                "traffic": traffic,
                "traffic_source": (f"profiles/{traffic_src}: committed ncu capture of this command, not measured by this "
                                   "run" if traffic_src else None),
                "peak_source": peak_src,
                "measured_in": ("the timed pass itself (CUDA events around every convolution launch)" if depth == 1 else
                                "single-stream pass of the same K steps (CUDA events around every launch)"),
                "ms_per_step_serial": ms_serial / args.steps,
                "launches_per_step": n_tc // max(args.steps, 1),
                "share_of_serial_step": tc_ms / ms_serial if world == 1 else None,
                "simt_conv_share_of_serial_step": simt_ms / ms_serial if world == 1 else None,
                "flops_per_step": tc_flops / max(args.steps, 1),
                # what the tensor pipe actually executes: every (256-row tile pair, offset) pass runs full M = 256 MMAs,
                # rows without that neighbour are zero-filled (not counted in `achieved`)
                "mma_executed": tc_exec / (tc_ms * 1e-3) / 1e12,
                "row_efficiency": tc_flops / tc_exec if tc_exec > 0 else None}

    if rank != 0:
        return
    total_frames = args.frames * world
    out = {
        "metric": "segmented-and-posed frames/s", "value": total_frames * args.steps / (ms_dev * 1e-3),
        "unit": "frames/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_dev / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": args.dtype, "data": "synthetic", "config": workload_config(args),
        "e2e": {"value": total_frames * args.steps / (ms_e2e * 1e-3), "unit": "frames/s",
                "h2d_bytes_per_step": int(B["N"] * (16 + 1)), "d2h_bytes_per_step": int(B["N"] + posed * 20 * 8),
                "ms_per_step": ms_e2e / args.steps},
        "gpu_launches": int(launches), "clocks": clocks, "roofline": roof,
        "frames_posed_per_step": int(posed) if world == 1 else int(np.nansum(allrec[:, 1])),
        "frames_confident_per_step": int(confident),
        "points_per_step": B["N"],
        "per_rank_ms_per_step": {"device": [float(x) / args.steps for x in per_rank[:, 0]],
                                 "e2e": [float(x) / args.steps for x in per_rank[:, 1]]},
    }
    if pred_crop is not None:
        out["pred_crop"] = pred_crop
    if stage_ms is not None:
        out["stage_ms"] = stage_ms
    out["accuracy"] = accuracy_record(args, frames, results, pose)
    if world == 1 and not args.no_cpu_baseline:
        out["cpu_baseline"], out["parity"] = cpu_baseline_and_parity(args, wb, frames)
    print(json.dumps(out))


def accuracy_record(args, frames, results, pose):
    """the evaluation record of app/test.py:78-290 over the frames of the last step (b200calib/evaluation.py), against
    the synthetic ground truth; written in full to --report, summarised in the JSON line. Random-init weights: the
    numbers say nothing about the method, they show that the record travels with every bench line."""
    from b200calib import evaluation as E
    inst = []
    init = {int(f): j for j, f in enumerate(pose["ok_frames"])} if pose.get("ee_pose_initial") is not None else {}
    for i, (f, r) in enumerate(zip(frames, results)):
        st = {"ee_points": f[0][f[2] == 2]}               # GT crop (the crop the timed steps use)
        # `pose` comes from the device-resident pass, `results` from the e2e pass: same frames, same weights
        if i in init:
            st["ee_pose_initial"] = pose["ee_pose_initial"][init[i]]
            if pose.get("kp_pose_initial") is not None and r.key_points_pose is not None:
                st["kp_pose_initial"] = pose["kp_pose_initial"][init[i]]
        rec = E.evaluate_frame(f[0], f[2], f[3], r, st)
        rec["position"] = f"p{i % 5}"                      # INFERENCE.CALIBRATION: frames grouped by robot position
        inst.append(rec)
    rep = E.aggregate(inst)
    if args.report:
        E.write_report(args.report, dict(report=rep, instances=inst))
    keys = ("segmentation_accuracy", "segmentation_precision", "segmentation_recall", "dist_position_nn",
            "dist_position_nn_icp", "angle_diff_nn", "angle_diff_nn_icp", "dist_position_kp_icp", "ADD_nn", "ADD_nn_icp")
    return {"frames": rep["frames"], "frames_confident": rep["frames_confident"], "units": "m / rad",
            "overall": {k: rep["overall"][k] for k in keys if k in rep["overall"]},
            "note": "random-init weights (checkpoints unpublished); full record: --report"}


def cpu_baseline_and_parity(args, wb, frames):
    """oracle (CPU restatement of the reference path) on ONE frame of the same workload, all host threads; the same
    frame then goes through the CUDA segmentation path in every compute mode and the per-point labels / logits are
    compared with the oracle's over ALL points (no margin mask)."""
    import torch
    import oracle.MinkowskiEngine as OME
    from oracle import pipeline as op
    from b200calib.synthetic import ee_surface_cloud
    from b200calib.pipeline import segment_points, normalize_colors_
    ME = wb.ME
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    seg, rot, kp = build_models(OME)
    cad = ee_surface_cloud(4096, SEED)
    p, c, l = frames[0][:3]
    t0 = time.perf_counter()
    ref = op.predict_frame(dict(seg=seg, rot=rot, kp=kp), cad, p, c, dict(seg_scale=args.scale,
                                                                           icp_enabled=not args.no_icp), gt_labels=l)
    dt = time.perf_counter() - t0
    base = {"value": 1.0 / dt, "unit": "frames/s", "cores": cores, "kind": "port",
            "sample": f"1 frame of the batch ({len(p)} points), batch 1 fp32, no warm-up, {dt:.1f} s"}
    raw = ref["segmentation_raw"]
    ref_lab, ref_log = raw["labels"], torch.from_numpy(raw["point_logits"]).double()
    parity = {"points": int(len(ref_lab)), "frame": "frame 0 of the batch, alone",
              "oracle": "oracle/pipeline.py predict_segmentation (fp32, CPU)",
              "oracle_label_histogram": np.bincount(ref_lab, minlength=3).tolist()}
    dpts, drgb = torch.from_numpy(p).to(wb.dev), torch.from_numpy(c).to(wb.dev)
    bidx = torch.zeros((len(p),), dtype=torch.float32, device=wb.dev)
    old = ME.get_compute_mode()
    try:
        for mode in ("f32", "tf32", "bf16"):
            ME.set_compute_dtype(mode)
            with torch.no_grad():
                lab, fld, out = segment_points(wb.seg, dpts, normalize_colors_(drgb), bidx, 1, args.scale)
                plog = out.slice(fld).F.double().cpu()
            parity[f"labels_mismatch_{'fp32' if mode == 'f32' else mode}"] = int((lab.cpu().numpy() != ref_lab).sum())
            parity[f"logits_rel_err_{'fp32' if mode == 'f32' else mode}"] = float((plog - ref_log).norm() / ref_log.norm())
    finally:
        ME.set_compute_dtype(old)
    return base, parity


# ---------------------------------------------------------------------------------------------- configs[3]: ICP 1k
def make_icp_problems(n, seed=SEED):
    """BASELINE.json configs[3] / SURVEY 8d C4: source = the CAD cloud (hand.pcd, 4480 points); per frame a random GT
    pose (t in [-0.5, 0.5]^2 x [0.8, 1.5] m), target = 2048..8192 points of the CAD side that faces the camera, sigma
    1.6 mm noise; init = GT pose with U[-0.03, 0.03] on all 7 entries, quaternion re-normalised
    (playground/play_ee_icp.py:113)."""
    from b200calib.transformation import get_transformation_matrix
    cad = load_cad().astype(np.float64)
    rng = np.random.default_rng(seed)
    tg, offs, T0, Tgt = [], [0], [], []
    for _ in range(n):
        q = rng.normal(size=4)
        q /= np.linalg.norm(q)
        t = np.array([rng.uniform(-0.5, 0.5), rng.uniform(-0.5, 0.5), rng.uniform(0.8, 1.5)])
        pose = np.concatenate((t, q))
        T = get_transformation_matrix(pose)
        pts = cad @ T[:3, :3].T + T[:3, 3]
        # visible side: points whose direction from the cloud centre faces the camera at the origin
        c = pts.mean(0)
        vis = ((pts - c) @ (-c / np.linalg.norm(c))) > -0.005
        cand = pts[vis]
        m = int(rng.integers(2048, 8193))
        sel = cand[rng.integers(0, len(cand), m)] + rng.normal(0, 0.0016, (m, 3))
        tg.append(sel.astype(np.float32))
        offs.append(offs[-1] + m)
        p0 = pose + rng.uniform(-0.03, 0.03, 7)
        p0[3:] /= np.linalg.norm(p0[3:])
        T0.append(get_transformation_matrix(p0))
        Tgt.append(T)
    return cad.astype(np.float32), np.concatenate(tg), np.asarray(offs, np.int32), np.stack(T0), np.stack(Tgt)


def icp_config(n):
    return {"workload": (f"point-to-point ICP refinement (utils/icp.py:50-81: Open3D registration_icp, max "
                         f"correspondence 0.1 m, <= 30 iterations, 1e-6) of {n} EE frames: source hand.pcd (4480 CAD "
                         f"points), targets 2048-8192 points, init jitter +-0.03; BASELINE.json configs[3]"),
            "frames": n, "source_points": 4480}


def run_icp1k(args, rank, world, local):
    if rank != 0:
        return
    import torch
    from b200calib.icp import icp_p2p_batched
    from oracle import geometry as og
    import MinkowskiEngine as ME
    n = 1000
    cad, tg, offs, T0, Tgt = make_icp_problems(n)
    dev = torch.device("cuda", local)
    torch.cuda.set_device(local)
    d_cad, d_tg, d_T0 = torch.from_numpy(cad).to(dev), torch.from_numpy(tg).to(dev), torch.from_numpy(T0).to(dev)
    h_tg, h_T0 = torch.from_numpy(tg).pin_memory(), torch.from_numpy(T0).pin_memory()
    for _ in range(args.warmup):
        T, st = icp_p2p_batched(d_cad, d_tg, offs, d_T0)
    torch.cuda.synchronize()
    sampler = ClockSampler(local)
    sampler.start()
    ME.reset_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        T, st = icp_p2p_batched(d_cad, d_tg, offs, d_T0)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    launches = ME.launch_count()
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2.record()
    for _ in range(args.steps):
        Te, ste = icp_p2p_batched(d_cad, h_tg.to(dev, non_blocking=True), offs, h_T0.to(dev, non_blocking=True))
        host_T = Te.cpu()
    e3.record()
    torch.cuda.synchronize()
    ms_e2e = e2.elapsed_time(e3) / args.steps
    clocks = sampler.stop()
    st_h, T_h = st.cpu().numpy(), T.cpu().numpy()
    iters = st_h[:, 2]
    evals = float((iters + 1).sum())                     # evaluation 0 + one per iteration
    S = cad.shape[0]
    queries = evals * S
    # pose error against the ground truth and parity against the CPU oracle on a sample of the frames
    err_t = np.linalg.norm(T_h[:, :3, 3] - Tgt[:, :3, 3], axis=1)
    sample = list(range(0, n, n // 20))[:20]
    t0 = time.perf_counter()
    dts, dang, dit = [], [], []
    for f in sample:
        To, fit, rmse, it = og.icp_point_to_point(cad, tg[offs[f]:offs[f + 1]], T0[f])
        dts.append(float(np.linalg.norm(To[:3, 3] - T_h[f, :3, 3])))
        dang.append(float(og.rotation_angle_deg(To[:3, :3], T_h[f, :3, :3])))
        dit.append(int(it) - int(iters[f]))
    cpu_s = (time.perf_counter() - t0) / len(sample)
    peaks = load_peaks()
    # algorithmic bytes: per evaluation S * 12 (source) + S * 8 (correspondence out), T * 12 once
    alg = evals * S * 20.0 + float(offs[-1]) * 12.0
    out = {"metric": "ICP-refined frames/s", "value": n / (ms * 1e-3), "unit": "frames/s", "n_gpus": 1,
           "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
           "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": icp_config(n),
           "e2e": {"value": n / (ms_e2e * 1e-3), "unit": "frames/s", "h2d_bytes_per_step": int(tg.nbytes + T0.nbytes),
                   "d2h_bytes_per_step": int(n * 16 * 8), "ms_per_step": ms_e2e},
           "gpu_launches": int(launches), "clocks": clocks,
           "icp": {"iterations_per_s": float(iters.sum()) / (ms * 1e-3), "iterations_mean": float(iters.mean()),
                   "iterations_max": int(iters.max()), "ns_per_nn_query": ms * 1e6 / queries,
                   "fitness_mean": float(st_h[:, 0].mean()), "rmse_mean": float(st_h[:, 1].mean()),
                   "translation_error_vs_gt_median_m": float(np.median(err_t))},
           "roofline": {"kernel": "k_icp_persistent (one cluster per frame, targets resident in shared memory)",
                        "bound": "hbm", "achieved": alg / (ms * 1e-3) / 1e9, "peak": peaks["hbm"], "unit": "GB/s",
                        "frac": alg / (ms * 1e-3) / 1e9 / peaks["hbm"], "traffic": None,
                        "note": "HBM only nominally: targets and cell lists live in shared memory, the kernel is bound "
                                "by the exact-NN search (fp64 distance tests per query), see ns_per_nn_query"},
           "parity": {"sample_frames": len(sample), "max_translation_diff_m": max(dts), "max_rotation_diff_deg": max(dang),
                      "iteration_count_diffs": dit, "tolerance": "1e-4 m / 0.01 deg (north star)"},
           "cpu_baseline": {"value": 1.0 / cpu_s, "unit": "frames/s", "cores": os.cpu_count() or 1, "kind": "port",
                            "sample": f"{len(sample)} of the {n} frames, oracle/geometry.py icp_point_to_point "
                                      f"(cKDTree + NumPy SVD), {cpu_s * 1e3:.0f} ms per frame"}}
    print(json.dumps(out))


def run_icp1k_reference(args, cores):
    from oracle import geometry as og
    n = 1000
    cad, tg, offs, T0, _ = make_icp_problems(n)
    per = 40                                              # frames per step: a bounded sample of the 1 000
    for i in range(args.warmup):
        og.icp_point_to_point(cad, tg[offs[i]:offs[i + 1]], T0[i])
    t0 = time.perf_counter()
    k = 0
    for s in range(args.steps):
        for j in range(per):
            f = (args.warmup + k) % n
            og.icp_point_to_point(cad, tg[offs[f]:offs[f + 1]], T0[f])
            k += 1
    dt = time.perf_counter() - t0
    fps = k / dt
    print(json.dumps({"impl": "reference", "metric": "ICP-refined frames/s", "value": fps, "unit": "frames/s",
                      "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
                      "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
                      "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": icp_config(n),
                      "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": cores, "kind": "port",
                                       "sample": f"{per} frames per step of the 1000 (cKDTree + NumPy SVD port of "
                                                 "Open3D's loop)"},
                      "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))


# ---------------------------------------------------------------------------------------------- configs[4]: sweep
def run_sweep(args, rank, world, local):
    """voxel-size sweep 0.02 -> 0.005 m on the configs[1] frames + a wide scene with >= 1 M voxels per frame at 5 mm,
    full pipeline, every rank its own frames (weak scaling), the CPU port timed beside it on rank 0."""
    wide_n = max(1, args.frames // 4)
    frames = make_workload(args.frames, rank, args.width, args.height)
    wide = make_workload(wide_n, rank, 2 * args.width, 2 * args.height, wide=True,
                         seeds=[SEED * 1000 + 900 + rank * 64 + i for i in range(wide_n)])
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(local)
    from b200calib import dist as bdist
    import MinkowskiEngine as ME
    bdist.init_from_env("nccl" if world > 1 else None)
    wb = Workbench(args, rank, world, local, frames)
    eng, dev = wb.eng, wb.dev
    rows = []
    for name, fr, scale in (("640x480 @ 20 mm", frames, 50.0), ("640x480 @ 10 mm", frames, 100.0),
                            ("640x480 @ 5 mm", frames, 200.0), ("wide 1280x960 @ 5 mm", wide, 200.0)):
        eng.cfg.seg_scale = scale
        B = wb.stage(fr)
        torch.cuda.reset_peak_memory_stats()

        def step():
            return eng.predict_device(B["d_pts"], B["d_rgb"], B["d_bidx"], B["offs"], gt_labels=B["d_lab"])
        census = ME.set_profile("census")
        step()
        ME.set_profile(None)
        v1 = max((c["V_out"] for c in census), default=0)
        for _ in range(max(args.warmup - 1, 0)):
            step()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        outs = list(eng.predict_stream(range(args.steps), depth=max(1, args.depth), fn=lambda i: step()))
        e1.record()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item()) / args.steps
        row = {"case": name, "frames_per_gpu": len(fr), "points_per_frame": int(B["N"] / len(fr)),
               "voxels_per_frame_stride1": int(v1 / len(fr)), "voxel_m": 1.0 / scale, "ms_per_step": ms,
               "frames_per_s": len(fr) * world / (ms * 1e-3), "frames_posed": int(len(outs[-1][1]["ok_frames"])),
               "peak_memory_gb": torch.cuda.max_memory_allocated() / 2 ** 30}
        if rank == 0 and not args.no_cpu_baseline:
            import oracle.MinkowskiEngine as OME
            from oracle import pipeline as op
            from b200calib.synthetic import ee_surface_cloud
            torch.set_num_threads(os.cpu_count() or 1)
            seg, rot, kp = build_models(OME)
            p, c, l = fr[0][:3]
            t0 = time.perf_counter()
            op.predict_frame(dict(seg=seg, rot=rot, kp=kp), ee_surface_cloud(4096, SEED), p, c,
                             dict(seg_scale=scale, icp_enabled=not args.no_icp), gt_labels=l)
            row["cpu_s_per_frame"] = time.perf_counter() - t0
            row["speedup_vs_cpu_port"] = row["frames_per_s"] * row["cpu_s_per_frame"]
        rows.append(row)
        del B
        torch.cuda.empty_cache()
    if rank != 0:
        return
    head = rows[2]
    print(json.dumps({"metric": "segmented-and-posed frames/s", "value": head["frames_per_s"], "unit": "frames/s",
                      "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": head["ms_per_step"],
                      "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": args.dtype,
                      "data": "synthetic",
                      "config": {"workload": "voxel-size sweep 0.02 -> 0.005 m + wide scene (>= 1 M voxels per frame), "
                                             "full pipeline; BASELINE.json configs[4]; `value` = the 5 mm row",
                                 "frames_per_gpu": args.frames, "batches_in_flight": args.depth},
                      "sweep": rows,
                      "cpu_baseline": {"value": 1.0 / head["cpu_s_per_frame"], "unit": "frames/s",
                                       "cores": os.cpu_count() or 1, "kind": "port",
                                       "sample": "1 frame per sweep row, batch 1 fp32"} if "cpu_s_per_frame" in head else None}))


# ---------------------------------------------------------------------------------------------- strong scaling
def run_strong(args, rank, world, local):
    """fixed workload of --strong-frames frames with uneven point counts (three render sizes), sharded over the ranks by
    greedy balancing on the pixel count (b200calib.dist.shard_frames), batches of <= --frames frames, gather of the
    pose records + calibrate() INSIDE the timed region."""
    F = args.strong_frames
    sizes_all = [((480, 360), (640, 480), (800, 600))[i % 3] for i in range(F)]
    weights = [w * h for w, h in sizes_all]
    from b200calib import dist as bdist
    mine = bdist.shard_frames(F, rank, world, weights=weights)
    frames = make_workload(len(mine), rank, args.width, args.height, seeds=[SEED * 1000 + 20000 + f for f in mine],
                           sizes=[sizes_all[f] for f in mine])
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(local)
    bdist.init_from_env("nccl" if world > 1 else None)
    wb = Workbench(args, rank, world, local, frames)
    eng, dev = wb.eng, wb.dev
    from b200calib.calibration import calibrate
    batches = [wb.stage(frames[i:i + args.frames]) for i in range(0, len(frames), args.frames)]
    ids = [mine[i:i + args.frames] for i in range(0, len(mine), args.frames)]
    ee2base = np.array([0.4, -0.1, 0.3, 0.9238795, 0.0, 0.3826834, 0.0])

    def one(B):
        labels, pose = eng.predict_device(B["d_pts"], B["d_rgb"], B["d_bidx"], B["offs"], gt_labels=B["d_lab"])
        return pose

    from b200calib.transformation import get_base2cam_pose

    class _Rec:   # the fields calibrate() reads (ResultDTO, app/dto.py:37-47)
        pass

    phase = {"predict": 0.0, "assemble": 0.0, "gather": 0.0, "calibrate": 0.0}   # host wall clock per phase, seconds

    def job():
        t_a = time.perf_counter()
        poses = list(eng.predict_stream(batches, depth=max(1, args.depth), fn=one))
        t_b = time.perf_counter()
        recs = []
        for bi, (B, pose) in enumerate(zip(batches, poses)):
            res = eng.assemble(np.zeros(B["N"], np.uint8), B["offs"], pose)
            recs.append(bdist.pack_records(ids[bi], res))
        t_c = time.perf_counter()
        allrec = bdist.gather_records(np.concatenate(recs) if recs else np.zeros((0, bdist.RECORD_WIDTH)))
        t_d = time.perf_counter()
        phase["predict"] += t_b - t_a
        phase["assemble"] += t_c - t_b
        phase["gather"] += t_d - t_c
        cal = None
        if rank == 0:   # InferenceEngine.calibrate over the gathered poses, frames grouped into 5 robot positions
            data = {}
            for row in allrec:
                if not row[1] > 0:
                    continue
                r = _Rec()
                r.is_confident = True
                r.ee_pose = row[3:10]
                r.base_pose = get_base2cam_pose(r.ee_pose, ee2base)
                r.key_points_pose = row[10:17] if np.isfinite(row[10]) else None
                r.key_points_base_pose = (get_base2cam_pose(r.key_points_pose, ee2base)
                                          if r.key_points_pose is not None else None)
                data.setdefault(int(row[0]) % 5, []).append(r)
            try:
                cal = calibrate(data) if data else None
            except ValueError:
                # the reference's calibrate (app/inference_engine.py:152-194) stacks base_pose with key_points_base_pose
                # and raises exactly like this when no frame of a position has a key-point pose (random-init weights)
                cal = None
        phase["calibrate"] += time.perf_counter() - t_d
        return allrec, cal

    for _ in range(max(1, args.warmup)):
        job()
    for k in phase:
        phase[k] = 0.0
    sampler = ClockSampler(local)
    sampler.start()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        allrec, cal = job()
    e1.record()
    torch.cuda.synchronize()
    ms_rank = e0.elapsed_time(e1) / args.steps
    if world > 1:
        dist.barrier()
    clocks = sampler.stop()
    t = torch.tensor([ms_rank, clocks["sm_mhz"] or 0.0, float(sum(len(f[0]) for f in frames)), float(len(frames))],
                     dtype=torch.float64, device=dev)
    if world > 1:
        tl = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(tl, t)
        per = torch.stack(tl).cpu().numpy()
    else:
        per = t.cpu().numpy()[None]
    if rank != 0:
        return
    ms = float(per[:, 0].max())
    print(json.dumps({"metric": "segmented-and-posed frames/s", "value": F / (ms * 1e-3), "unit": "frames/s",
                      "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
                      "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": args.dtype,
                      "data": "synthetic",
                      "config": {"workload": f"strong scaling: {F} frames in total (480x360 / 640x480 / 800x600 renders), "
                                             f"greedy balancing by pixel count, batches of <= {args.frames}, gather of "
                                             f"the pose records inside the timed region",
                                 "frames_total": F, "batches_in_flight": args.depth},
                      "frames_posed": int(np.nansum(allrec[:, 1])),
                      "host_phase_ms_rank0": {k: round(v * 1e3 / args.steps, 2) for k, v in phase.items()},
                      "calibration_pose": None if cal is None or cal.pose_camera_link is None
                      else [float(x) for x in cal.pose_camera_link],
                      "per_rank": [{"rank": r, "ms_per_step": float(per[r, 0]), "sm_mhz": float(per[r, 1]),
                                    "points": int(per[r, 2]), "frames": int(per[r, 3])} for r in range(world)],
                      "clocks": clocks}))


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_b200(args, rank, world, local)
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            dist.destroy_process_group()


if __name__ == "__main__":
    main()
