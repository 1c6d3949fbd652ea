#!/usr/bin/env python
"""Collects the JSON lines the bench printed in the round's gpurun calls into profiles/r02_bench_lines.md.
   python profiles/collect_bench_lines.py gpurun_out"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ITEMS = [  # (log, command, note)
    ("bench_default.log", "python bench.py   (tools/gpu_run.sh, the round-end validation script)", "final commit, default flags, 1 GPU"),
    ("r2_bench17.log", "python bench.py --stages --no-cpu-baseline", "final build, 1 GPU, 1413 MHz box"),
    ("r2_bench16.log", "python bench.py --stages --no-cpu-baseline", "final build (+ ICP pre-filter), 1 GPU, 1342 MHz box"),
    ("r2_bench15.log", "python bench.py --stages --no-cpu-baseline", "final build (+ 2x2x2-group table lines in K1), 1 GPU, 1432 MHz box"),
    ("r2_bench13.log", "python bench.py --stages --conv-table ...", "default run, final build (TMA operand path), 1 GPU"),
    ("r2_bench13_again.log", "python bench.py --no-cpu-baseline", "same build, same box, later in the call"),
    ("r2_bench13_cpasync.log", "python bench.py --tc-path cpasync --no-cpu-baseline", "same box: cp.async operand path"),
    ("r2_bench13_depth2.log", "python bench.py --depth 2 --no-cpu-baseline", "same box: two batches in flight"),
    ("r2_bench11.log", "python bench.py --stages --crop both ...", "run 11 (cp.async path default), other box; carries `pred_crop`"),
    ("r2_bench26_tf32.log", "python bench.py --dtype tf32 --steps 5 --no-cpu-baseline", "tf32 tensor-core mode, final build"),
    ("r2_bench10_tf32.log", "python bench.py --dtype tf32 --steps 5", "tf32 tensor-core mode, run 10 (cp.async path)"),
    ("r2_bench8_vote.log", "python bench.py --vote --no-cpu-baseline --stages --steps 4", "with the vote stage"),
    ("r2_icp1k17.log", "python bench.py --config icp1k", "BASELINE configs[3], final build (fp32 pre-filter + four candidates per step in the exact NN search)"),
    ("r2_icp1k16.log", "python bench.py --config icp1k", "BASELINE configs[3], fp32 pre-filter only"),
    ("r2_icp1k10.log", "python bench.py --config icp1k", "BASELINE configs[3], before the pre-filter"),
    ("r2_sweep27.log", "python bench.py --config sweep --steps 3 --warmup 2 --no-cpu-baseline", "BASELINE configs[4], 1 GPU, final build"),
    ("r2_sweep7.log", "python bench.py --config sweep", "BASELINE configs[4], 1 GPU, run 7: CPU port timed beside every row (GPU rows of this run carry the allocator artefact fixed later)"),
    ("r2_sweep_8gpu.log", "torchrun --nproc-per-node 8 bench.py --gpus 8 --config sweep --steps 3 --warmup 2 --no-cpu-baseline", "configs[4] at 8 GPUs"),
    ("r2_bench_8gpu.log", "torchrun --nproc-per-node 8 bench.py --gpus 8 --steps 10 --warmup 3", "weak scaling, 32 frames per GPU"),
    ("r2_strong256_1gpu.log", "python bench.py --strong-frames 256 --steps 2 --warmup 1", "strong scaling reference: 256 uneven frames on 1 GPU (same 8-GPU box)"),
    ("r2_strong256_8gpu.log", "torchrun --nproc-per-node 8 bench.py --gpus 8 --strong-frames 256 --steps 3 --warmup 1", "strong scaling: the same 256 frames on 8 GPUs, gather + calibrate inside the timed region"),
    ("r2_strong1000_8gpu.log", "torchrun --nproc-per-node 8 bench.py --gpus 8 --strong-frames 1000 --steps 1 --warmup 1", "strong scaling: 1000 frames on 8 GPUs"),
    ("r2_ref7.log", "python bench.py --impl reference --steps 2 --warmup 1", "reference arm: CPU oracle port on the box's 16 host cores"),
]


def main(d):
    L = ["# Bench JSON lines of round 2 (as printed; one gpurun call = one box, boxes differ in their power-capped clock)\n"]
    for log, cmd, note in ITEMS:
        path = os.path.join(d, log)
        if not os.path.exists(path):
            continue
        lines = [x for x in open(path) if x.startswith("{")]
        if not lines:
            continue
        j = json.loads(lines[-1])
        head = f"value {j.get('value'):.1f} {j.get('unit')}, {j.get('ms_per_step', 0):.1f} ms/step, n_gpus {j.get('n_gpus')}"
        if j.get("e2e"):
            head += f", e2e {j['e2e']['value']:.1f}"
        if j.get("roofline") and j["roofline"].get("frac") is not None:
            head += f", roofline.frac {j['roofline']['frac']:.3f}"
        if j.get("clocks"):
            head += f", SM {j['clocks'].get('sm_mhz')} MHz {j['clocks'].get('reasons')}"
        L.append(f"## {note}\n\n`{cmd}` ({log})\n\n{head}\n\n```json\n{lines[-1].strip()}\n```\n")
    open(os.path.join(HERE, "r02_bench_lines.md"), "w").write("\n".join(L))


if __name__ == "__main__":
    main(sys.argv[1])
