#!/usr/bin/env python
"""bench.py — segmented-and-posed frames/s of the markerless-calibration inference hot path on B200.

One "step" = one pass of the whole per-batch pipeline over `--frames` synthetic 640x480 Kinect-shaped frames per
GPU (BASELINE.json configs[1] + [2]): voxelise (K1) -> kernel maps (K2/K3) -> MinkUNet18D segmentation forward in
bf16 on tcgen05 (K4) -> head + per-point arg-max (K5) -> largest EE cluster (K7) -> RobotNetEncode rotation (K1-K4,
K6) -> "magic" translation -> key-point network + per-class reduction (K8) -> batched Kabsch (K9) -> ICP x2 (K10).
Random-init weights of the named architectures, synthetic data (dataset and checkpoints are unpublished).

  python bench.py --gpus N --steps K --warmup W            # B200 arm (under torchrun for N > 1)
  python bench.py --impl reference --gpus N ...            # CPU oracle of the reference path on the host cores

Prints ONE JSON line (rank 0). `value` = frames/s with inputs resident in HBM; `e2e` = the same through the public
API with pinned host buffers (H2D of the inputs, D2H of labels and poses inside the timed region).
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
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--frames", type=int, default=32, help="frames per GPU per step (batch)")
    ap.add_argument("--width", type=int, default=640)
    ap.add_argument("--height", type=int, default=480)
    ap.add_argument("--scale", type=float, default=200.0, help="voxels per metre (200 = 5 mm, production)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-icp", action="store_true")
    ap.add_argument("--stages", action="store_true", help="one extra (untimed) step with synchronised per-stage times")
    ap.add_argument("--dtype", default="bf16", choices=["bf16", "f32"])
    ap.add_argument("--kp-backbone", default="minkunet", choices=["minkunet", "pointnet2"],
                    help="key-point network: 6-class MinkUNet18D (default; both arms) or the reference's default "
                         "PointNet2SSG branch on GPU-native FPS / ball query / 3-NN (B200 arm only)")
    ap.add_argument("--mask-block", type=int, default=None, help="rows per locality block of the K3b mask sort (0 = global)")
    ap.add_argument("--mask-morton", action="store_true", help="A/B: Morton order inside a mask group of the k3 maps")
    ap.add_argument("--mask-two-level", action="store_true", help="A/B: two-level K3b mask-sort keys (default: one-level)")
    ap.add_argument("--cprofile", default=None, help="one extra (untimed) step under cProfile: host-side launch cost (text)")
    ap.add_argument("--torch-profile", default=None,
                    help="one extra (untimed) step under torch.profiler: per-kernel device times + GPU busy fraction (text)")
    ap.add_argument("--conv-table", default=None, help="write the per-convolution census + CUDA-event times here (JSON)")
    return ap.parse_args()


def _gen_frame(a):
    from b200calib.synthetic import make_frame
    seed, w, h = a
    f = make_frame(seed, width=w, height=h)
    rgb = (np.round(f["rgb"] * 255.0) / 255.0).astype(np.float32)   # 8-bit colours, as a camera delivers them
    return f["points"], rgb, f["labels"].astype(np.uint8)


def make_workload(n_frames, rank, width, height):
    """n_frames distinct seeded frames for this rank (generated before CUDA is touched, in worker processes)."""
    import multiprocessing as mp
    seeds = [(SEED * 1000 + rank * 4096 + i, width, height) for i in range(n_frames)]
    nproc = max(1, min(8, (os.cpu_count() or 8) // max(1, int(os.environ.get("WORLD_SIZE", "1")))))
    if nproc > 1 and n_frames > 1:
        with mp.get_context("fork").Pool(nproc) as pool:
            return pool.map(_gen_frame, seeds)
    return [_gen_frame(s) for s in seeds]


def load_traffic(frames):
    """dram__bytes_read.sum + dram__bytes_write.sum per k_spconv_tc launch (mean over the 121 launches of one step), from
    the committed ncu capture of this command at 32 frames (profiles/r01_k_spconv_tc_dram_traffic.json); None for
    another batch size or when the file is absent."""
    path = os.path.join(ROOT, "profiles", "r01_k_spconv_tc_dram_traffic.json")
    try:
        with open(path) as fp:
            t = json.load(fp)
        return float(t["dram_bytes_per_launch"]) if int(t.get("frames", -1)) == int(frames) else None
    except (OSError, ValueError, KeyError):
        return None


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], tf_burst=d["bf16_tflops"], tf_sustained=d.get("bf16_tflops_sustained",
                                                                                 d["bf16_tflops"]), src="measured")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sustained=1400.0, src="fallback")


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


# ---------------------------------------------------------------------------------------------- reference arm
def run_reference(args, rank, world):
    """the reference's own per-frame CPU path, restated (oracle/pipeline.py; the real one needs MinkowskiEngine 0.5.4
    + Open3D, absent here), on all host threads. Each step = ONE frame of the workload (a bounded sample)."""
    if rank != 0:
        return
    import torch
    import oracle.MinkowskiEngine as OME
    from oracle import pipeline as op
    from b200calib.synthetic import ee_surface_cloud
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    seg, rot, kp = build_models(OME)
    nfr = min(args.frames, args.warmup + args.steps)
    frames = make_workload(max(1, min(nfr, 2)), 0, args.width, args.height)
    cad = ee_surface_cloud(4096, SEED)
    cfg = dict(seg_scale=args.scale, icp_enabled=not args.no_icp)
    models = dict(seg=seg, rot=rot, kp=kp)

    def one(i):
        p, c, l = frames[i % len(frames)]
        return op.predict_frame(models, cad, p, c, cfg, gt_labels=l)

    for i in range(min(args.warmup, 1)):  # one warm-up frame is ~30 s of CPU work; more would not change the number
        one(i)
    t0 = time.perf_counter()
    for i in range(args.steps):
        one(i)
    dt = time.perf_counter() - t0
    fps = args.steps / dt
    sample = f"{args.steps} frames, one 640x480 frame per step, batch 1 fp32 (the reference's mode)"
    print(json.dumps({
        "impl": "reference", "metric": "segmented-and-posed frames/s", "value": fps, "unit": "frames/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": min(args.warmup, 1), "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args),
        "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def workload_config(args):
    kpn = "6-key-point MinkUNet18D" if args.kp_backbone == "minkunet" else "6-key-point PointNet2SSG on 2048 points"
    return {"workload": (f"segmentation forward (MinkUNet18D, 3 classes) + EE pose (RobotNetEncode rotation, magic "
                         f"translation, {kpn} + Kabsch, ICP x2) on {args.frames} synthetic "
                         f"{args.width}x{args.height} Kinect-shaped frames per GPU, voxel {1.0 / args.scale * 1000:.1f} mm "
                         f"(seg/rot), 1.25 mm (key points); BASELINE.json configs[1]+[2]"),
            "frames_per_gpu": args.frames, "points_per_frame": "~3.0e5", "voxel_m": 1.0 / args.scale,
            "ee_crop": "ground-truth labels (random-init weights give no usable EE prediction)",
            "l2": "inputs and activations exceed L2 (126 MB) every step; no explicit flush",
            "parallelism": "frames sharded across GPUs, one final all_gather of pose records"}


# ---------------------------------------------------------------------------------------------- B200 arm
def run_b200(args, rank, world, local):
    frames = make_workload(args.frames, rank, args.width, args.height)   # before CUDA init (fork pool)
    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py (B200 arm) needs a CUDA device; there is no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    import MinkowskiEngine as ME
    from b200calib import dist as bdist
    from b200calib.pipeline import BatchedInferenceEngine, PipelineConfig
    from b200calib.synthetic import ee_surface_cloud
    bdist.init_from_env("nccl" if world > 1 else None)
    ME.set_compute_dtype(torch.bfloat16 if args.dtype == "bf16" else torch.float32)
    if args.mask_block is not None:
        ME.set_mask_sort_block(args.mask_block)
    if args.mask_morton:
        ME.set_mask_sort_morton(True)
    if args.mask_two_level:
        ME.set_mask_sort_two_level(True)
    seg, rot, kp = [m.to(dev) for m in build_models(ME, kp_backbone=args.kp_backbone)]
    cad = torch.from_numpy(ee_surface_cloud(4096, SEED)).to(dev)
    cfg = PipelineConfig(seg_scale=args.scale, icp_enabled=not args.no_icp)
    eng = BatchedInferenceEngine(seg, rot, kp, cad_points=cad, config=cfg)

    # pinned host staging of this rank's batch (the "host buffers" of the e2e number)
    counts = [len(f[0]) for f in frames]
    offs = np.zeros(len(frames) + 1, dtype=np.int64)
    np.cumsum(counts, out=offs[1:])
    N = int(offs[-1])
    h_pts = torch.from_numpy(np.concatenate([f[0] for f in frames])).pin_memory()
    h_rgb = torch.from_numpy(np.concatenate([f[1] for f in frames])).pin_memory()
    h_bidx = torch.from_numpy(np.repeat(np.arange(len(frames), dtype=np.float32), counts)).pin_memory()
    h_lab = torch.from_numpy(np.concatenate([f[2] for f in frames])).pin_memory()
    h_seg = torch.empty((N,), dtype=torch.uint8).pin_memory()
    d_pts, d_rgb, d_bidx, d_lab = [t.to(dev) for t in (h_pts, h_rgb, h_bidx, h_lab)]
    torch.cuda.synchronize()

    def step_device():
        return eng.predict_device(d_pts, d_rgb, d_bidx, offs, gt_labels=d_lab)

    # e2e host buffers: PointCloud2-style records (x, y, z, PCL-packed rgb; 16 B per point), the wire format the
    # reference's live source delivers (app/freenect_data_engine.py:74-81); b200calib.ingest unpacks / normalises /
    # ROI-filters them on the device
    from b200calib.ingest import ingest_clouds, pack_xyzrgb
    h_rec = torch.from_numpy(np.concatenate([pack_xyzrgb(f[0], np.round(f[1] * 255.0)) for f in frames])).pin_memory()
    offs32 = offs.astype(np.int32)

    def step_e2e():
        r = h_rec.to(dev, non_blocking=True)
        g = h_lab.to(dev, non_blocking=True)
        p, c, b, o = ingest_clouds(r, offs32)      # synthetic frames have no invalid pixels: nothing is dropped
        labels, pose = eng.predict_device(p, c, b, o, gt_labels=g, rgb_normalized=True)
        h_seg.copy_(labels, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return eng.assemble(h_seg.numpy(), o, pose)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- warm-up (also: census of the algorithmic work of every convolution launch, once)
    recs_census = ME.set_profile("census")
    step_device()
    ME.set_profile(None)
    for _ in range(max(args.warmup - 1, 0)):
        step_device()
    barrier()

    # ---- timed: device-resident
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
    ms_dev = e0.elapsed_time(e1)
    launches = ME.launch_count()
    ME.set_profile(None)
    clocks = sampler.stop() if rank == 0 else None

    # ---- timed: end to end through the public API with pinned host buffers
    step_e2e()
    barrier()
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2.record()
    for _ in range(args.steps):
        results = step_e2e()
    e3.record()
    barrier()
    ms_e2e = e2.elapsed_time(e3)
    posed = sum(1 for r in results if r.ee_pose is not None)

    stage_ms = None
    if args.stages:
        eng.stage_times = {}
        step_device()
        stage_ms = {k: round(v, 3) for k, v in eng.stage_times.items()}
        eng.stage_times = None

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
    t = torch.tensor([ms_dev, ms_e2e], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_dev, ms_e2e = t.tolist()

    # ---- roofline of the dominant kernel (k_spconv_tc): FLOPs from the census / CUDA-event durations
    per_step = len(recs_census)
    tc_flops = tc_ms = simt_ms = tc_exec = 0.0
    all_ms = 0.0
    n_tc = 0
    torch.cuda.synchronize()
    for i, ev in enumerate(recs_ev):
        c = recs_census[i % per_step]
        ms = ev[0].elapsed_time(ev[1])
        all_ms += ms
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
        roof = {"kernel": "k_spconv_tc (tcgen05 gather-GEMM sparse convolution)", "bound": "tensor", "achieved": ach,
                "peak": peaks["tf_sustained"], "unit": "TFLOP/s", "frac": ach / peaks["tf_sustained"],
                "traffic": load_traffic(args.frames),
                "peak_source": f"{peaks['src']} bf16 sustained (burst {peaks['tf_burst']})",
                "launches_per_step": n_tc // max(args.steps, 1),
                "share_of_step": tc_ms / ms_dev if world == 1 else None,
                "simt_conv_share_of_step": simt_ms / ms_dev if world == 1 else None,
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
                "h2d_bytes_per_step": int(N * (16 + 1)), "d2h_bytes_per_step": int(N + posed * 20 * 8),
                "ms_per_step": ms_e2e / args.steps},
        "gpu_launches": int(launches), "clocks": clocks, "roofline": roof,
        "frames_posed_per_step": int(posed) if world == 1 else int(np.nansum(allrec[:, 1])),
        "points_per_step": N,
    }
    if stage_ms is not None:
        out["stage_ms"] = stage_ms
    if world == 1 and not args.no_cpu_baseline:
        out["cpu_baseline"] = cpu_baseline(args, frames)
    print(json.dumps(out))


def cpu_baseline(args, frames):
    """oracle (CPU restatement of the reference path) on ONE frame of the same workload, all host threads."""
    import torch
    import oracle.MinkowskiEngine as OME
    from oracle import pipeline as op
    from b200calib.synthetic import ee_surface_cloud
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    seg, rot, kp = build_models(OME)
    cad = ee_surface_cloud(4096, SEED)
    p, c, l = frames[0]
    t0 = time.perf_counter()
    op.predict_frame(dict(seg=seg, rot=rot, kp=kp), cad, p, c, dict(seg_scale=args.scale,
                                                                     icp_enabled=not args.no_icp), gt_labels=l)
    dt = time.perf_counter() - t0
    return {"value": 1.0 / dt, "unit": "frames/s", "cores": cores, "kind": "port",
            "sample": f"1 frame of the batch ({len(p)} points), batch 1 fp32, no warm-up, {dt:.1f} s"}


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
