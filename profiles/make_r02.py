#!/usr/bin/env python
"""Builds the round-2 evidence files of profiles/ from the ncu exports of one gpurun call (tools/gpu_run10.sh):
   python profiles/make_r02.py gpurun_out 10
Inputs (all written on the GPU box by `ncu -i rep --page ... --csv`; the .ncu-rep files are too large to bring back):
   r2_launches<N>.csv, r2_prof<N>_tc_raw.csv, r2_prof<N>_tc_details.csv, r2_prof<N>_misc_raw.csv, r2_tc_dram<N>.csv"""
import csv
import json
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
HBM = 6557.1  # GB/s, MEASURED_PEAKS.json (copy)


def raw_table(path):
    rows = list(csv.reader(line for line in open(path, newline="") if not line.startswith("==")))
    hdr, units = rows[0], rows[1]
    out = []
    for r in rows[2:]:
        if len(r) == len(hdr):
            out.append({h: v for h, v in zip(hdr, r)})
    return out, dict(zip(hdr, units))


def scale(v, unit, to):
    f = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12, "ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0,
         "usecond": 1e-6, "msecond": 1e-3, "nsecond": 1e-9, "second": 1.0}
    return float(v.replace(",", "")) * f[unit.split("/")[0]] / f[to]


def tc(out, d, n):
    rows, units = raw_table(os.path.join(d, f"r2_prof{n}_tc_raw.csv"))
    g = lambda r, k, to: scale(r[k], units[k], to)
    L = ["# `ncu --set full` of the last nine `k_spconv_tc` launches of the segmentation forward (32 frames, one warm step)\n",
         "Command: `ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:k_spconv_tc -s 40 "
         "-c 9 python bench.py --frames 32 --steps 1 --warmup 1 --no-cpu-baseline --profile-range` (run %s; the report is "
         "exported to CSV on the GPU box, `tools/gpu_run%s.sh`). V = 3.54 M rows at tensor stride 2, 8.92 M at stride 1.\n" % (n, n),
         "| launch | kernel | what | ms | tensor pipe active % | issue slots busy % | DRAM read GB | DRAM write GB | DRAM GB/s (% of 6557) | L2 hit % | registers | dyn smem KB |",
         "|---:|---|---|---:|---:|---:|---:|---:|---:|---:|---:|---:|"]
    what = ["K27 384->384 stride 2 (last of block7)", "K8 384->384 transposed, up to stride 1", "K27 416->384 stride 1 (cat)",
            "K1 416->384 residual projection", "K27 384->384 stride 1", "K27 384->384 stride 1", "K27 384->384 stride 1",
            "K1 384->256", "fused head 256->1024->3 + argmax"]
    for i, r in enumerate(rows):
        ms = g(r, "gpu__time_duration.sum", "ms")
        rd, wr = g(r, "dram__bytes_read.sum", "Gbyte"), g(r, "dram__bytes_write.sum", "Gbyte")
        bw = (rd + wr) / ms * 1e3
        L.append("| %s | `%s` | %s | %.2f | %.1f | %.1f | %.2f | %.2f | %.0f (%.0f %%) | %.1f | %s | %.0f |" % (
            r["ID"], r["Kernel Name"].replace("void ", "").replace("(TcParams)", ""), what[i] if i < len(what) else "", ms,
            float(r["sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active"]),
            float(r["smsp__issue_active.avg.pct_of_peak_sustained_active"]), rd, wr, bw, 100 * bw / HBM,
            float(r["lts__t_sector_hit_rate.pct"]), r.get("launch__registers_per_thread", "?"),
            g(r, "launch__shared_mem_per_block_dynamic", "Kbyte") if "launch__shared_mem_per_block_dynamic" in r else 0))
    L.append("")
    L.append("Per-launch section tables (`--page details`):\n")
    open(out, "w").write("\n".join(L))
    det = os.path.join(d, f"r2_prof{n}_tc_details.csv")
    md = subprocess.run([sys.executable, os.path.join(HERE, "ncu_details_to_md.py")], stdin=open(det), capture_output=True,
                        text=True).stdout
    open(out, "a").write(md)


ALG = {  # algorithmic bytes of the stride-1 instance (N = 9.48 M points, V = 8.92 M voxels, 32 frames), DESIGN.md section 3
    "k_hash_insert": ("N*16 + N*4", lambda N, V: N * 20),
    "k_first_flags": ("N*4 + N*4", lambda N, V: N * 8),
    "k_assign_rows": ("N*(16+4+4) + V*(16+4)", lambda N, V: N * 24 + V * 20),
    "k_inverse_accumulate": ("N*(4+12) + N*4", lambda N, V: N * 20),
    "k_quantize_float": ("N*16 + N*16", lambda N, V: N * 32),
    "k_kernel_map_k3_blocks": ("V*16 + V*27*4 + V*4", lambda N, V: V * 128),
    "k_block_rows": ("V*16 + V*4 + B*256", lambda N, V: V * 20 + V / 8 * 256),
    "k_stride_kernel_maps": ("(V_out+V_in)*32", lambda N, V: (V + V / 2.5) * 32),
    "k_color_apply": ("N*12 * 2", lambda N, V: N * 24),
    "k_color_stats": ("N*12", lambda N, V: N * 12),
}


def misc(out, d, n, N=9.48e6, V=8.92e6):
    rows, units = raw_table(os.path.join(d, f"r2_prof{n}_misc_raw.csv"))
    g = lambda r, k, to: scale(r[k], units[k], to)
    L = ["# `ncu --set full` of the coordinate / pre-processing kernels (32 frames, one warm step)\n",
         "Command: `ncu --profile-from-start off --set full --clock-control none -k regex:<names> -c 40 python bench.py "
         "--frames 32 --steps 1 --warmup 1 --no-cpu-baseline --profile-range` (run %s). First 40 matching launches = colour "
         "normalisation, K1 of the segmentation field (N = 9.48 M points -> V = 8.92 M voxels) and the K2 / K3 passes of the "
         "coordinate levels below it. `alg. GB/s` = algorithmic bytes (DESIGN.md section 3) of the stride-1 instance / duration; "
         "`traffic ratio` = measured DRAM bytes / algorithmic bytes.\n" % n,
         "| launch | kernel | us | DRAM read MB | DRAM write MB | DRAM GB/s (% of 6557) | alg. bytes | alg. GB/s (frac) | traffic ratio | L2 hit % | warps active % | threads / inst |",
         "|---:|---|---:|---:|---:|---:|---|---:|---:|---:|---:|---:|"]
    seen = set()
    for r in rows:
        name = r["Kernel Name"].split("(")[0].replace("void ", "")
        us = g(r, "gpu__time_duration.sum", "us")
        rd, wr = g(r, "dram__bytes_read.sum", "Mbyte"), g(r, "dram__bytes_write.sum", "Mbyte")
        bw = (rd + wr) / us * 1e3  # MB/us = TB/s -> GB/s
        alg = ""
        algbw = ""
        ratio = ""
        if name in ALG and name not in seen:   # first instance = the stride-1 (largest) one
            seen.add(name)
            b = ALG[name][1](N, V)
            alg = "%s = %.0f MB" % (ALG[name][0], b / 1e6)
            algbw = "%.0f (%.2f)" % (b / us * 1e-3, b / us * 1e-3 / HBM)
            ratio = "%.1f" % ((rd + wr) * 1e6 / b)
        L.append("| %s | `%s` | %.1f | %.0f | %.0f | %.0f (%.0f %%) | %s | %s | %s | %.1f | %.1f | %.1f |" % (
            r["ID"], name, us, rd, wr, bw, 100 * bw / HBM, alg, algbw, ratio, float(r["lts__t_sector_hit_rate.pct"]),
            float(r["sm__warps_active.avg.pct_of_peak_sustained_active"]),
            float(r["smsp__thread_inst_executed_per_inst_executed.ratio"])))
    open(out, "w").write("\n".join(L) + "\n")


def dram(out_md, out_json, d, n):
    path = os.path.join(d, f"r2_tc_dram{n}.csv")
    rows = list(csv.DictReader(line for line in open(path, newline="") if not line.startswith("==")))
    per = {}
    for r in rows:
        per.setdefault(r["ID"], {"name": r["Kernel Name"]})[r["Metric Name"]] = scale(
            r["Metric Value"], r["Metric Unit"], "byte" if "bytes" in r["Metric Name"] else "ms")
    L = []
    rd = sum(v["dram__bytes_read.sum"] for v in per.values())
    wr = sum(v["dram__bytes_write.sum"] for v in per.values())
    ms = sum(v["gpu__time_duration.sum"] for v in per.values())
    cmd = ("ncu --profile-from-start off --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum "
           "--clock-control none -k regex:k_spconv_tc --csv python bench.py --frames 32 --steps 1 --warmup 1 "
           "--no-cpu-baseline --profile-range")
    L.append("# DRAM traffic of k_spconv_tc over one warm 32-frame step (round 2, run %s)\n" % n)
    L.append("`%s` (%d launches = one step: segmentation + rotation + key-point networks, fused head included).\n" % (cmd, len(per)))
    L.append("Per step: **%.1f GB read + %.1f GB written = %.1f GB** in %.1f ms of kernel time under ncu = %.2f TB/s; mean "
             "%.2f GB per launch (= `roofline.traffic`). Round 1: 417 + 88 = 505 GB (the fused head removed the 18 GB write of "
             "the 256->1024 hidden tensor and its re-read).\n" % (rd / 1e9, wr / 1e9, (rd + wr) / 1e9, ms, (rd + wr) / ms / 1e9,
                                                                 (rd + wr) / len(per) / 1e9))
    L.append("| launch | kernel | ms (ncu) | DRAM read GB | DRAM write GB | GB/s |")
    L.append("|---:|---|---:|---:|---:|---:|")
    for k, v in per.items():
        if v["gpu__time_duration.sum"] >= 1.0:
            L.append("| %s | `%s` | %.2f | %.2f | %.2f | %.0f |" % (
                k, v["name"].replace("void ", "").replace("(TcParams)", ""), v["gpu__time_duration.sum"],
                v["dram__bytes_read.sum"] / 1e9, v["dram__bytes_write.sum"] / 1e9,
                (v["dram__bytes_read.sum"] + v["dram__bytes_write.sum"]) / v["gpu__time_duration.sum"] / 1e6))
    L.append("\n(launches shorter than 1 ms omitted from the table, included in the sums)")
    open(out_md, "w").write("\n".join(L) + "\n")
    json.dump({"frames": 32, "launches": len(per), "dram_read_bytes_per_step": rd, "dram_write_bytes_per_step": wr,
               "dram_bytes_per_launch": (rd + wr) / len(per), "kernel_ms_per_step_under_ncu": ms, "command": cmd,
               "run": int(n)}, open(out_json, "w"), indent=1)


def main():
    d, n = sys.argv[1], sys.argv[2]
    p = lambda f: os.path.join(HERE, f)
    if len(sys.argv) > 3 and sys.argv[3] == "misc":   # only the coordinate-kernel table, from a later run
        misc(p("r02_ncu_coords_kernels.md"), d, n)
        return
    with open(p("r02_launches_step32.md"), "w") as fp:
        fp.write("# Launch list of ONE warm 32-frame step (round 2, run %s)\n\n`ncu --profile-from-start off --metrics "
                 "gpu__time_duration.sum --clock-control none --csv python bench.py --frames 32 --steps 1 --warmup 1 "
                 "--no-cpu-baseline --profile-range` (cudaProfilerStart/Stop around one step after the warm-up, so the "
                 "census pass and its torch reductions are not in the list).\n\n" % n)
        fp.write(subprocess.run([sys.executable, p("summarize_launches.py"), os.path.join(d, f"r2_launches{n}.csv")],
                                capture_output=True, text=True).stdout)
    tc(p("r02_ncu_k_spconv_tc.md"), d, n)
    misc(p("r02_ncu_coords_kernels_run%s.md" % n), d, n)
    dram(p("r02_k_spconv_tc_dram_traffic.md"), p("r02_k_spconv_tc_dram_traffic.json"), d, n)


if __name__ == "__main__":
    main()
