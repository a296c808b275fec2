#!/usr/bin/env python
"""profiles/ summaries out of an ncu launch list (tools/profile_r2.sh): one train step's launches with duration, tensor-pipe %
and DRAM traffic; the per-step DRAM bytes of the tensor-core kernels (bench.py's roofline.traffic) as JSON.
    python tools/summarize_launches.py gpurun_out/r2_launches.csv r2"""
import collections
import csv
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main(fn, tag):
    lines = [l for l in open(fn) if not l.startswith("==")]
    launch = collections.OrderedDict()
    for r in csv.DictReader(lines):
        d = launch.setdefault(r["ID"], {"name": r["Kernel Name"], "grid": r["Grid Size"]})
        v = float(r["Metric Value"].replace(",", ""))
        u = r.get("Metric Unit", "")
        if r["Metric Name"].startswith("gpu__time"):
            v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(u, 1.0)
        elif "bytes" in r["Metric Name"]:
            v *= {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
        d[r["Metric Name"]] = v
    rows = list(launch.values())
    for r in rows:
        r["short"] = re.sub(r"\(.*", "", r["name"].replace("void ", "")).replace("mfvae::", "")
    starts = [i for i, r in enumerate(rows) if r["short"].startswith("stage_kernel")]
    step = rows[starts[0]:starts[1]]
    out = [f"one train step of tools/one_step.py (cfg2, B=4096, bf16) under ncu --metrics gpu__time_duration.sum,dram__bytes_*.sum,"
           f"sm__pipe_tensor_cycles_active...,gpu__dram_throughput... --clock-control none (serialised, cold cache)",
           "columns: kernel | grid | duration us | tensor pipe % of peak | DRAM % of peak | DRAM MB (read + write)"]
    tt = tb = 0.0
    agg = collections.OrderedDict()
    for r in step:
        t = r["gpu__time_duration.sum"]; b = r["dram__bytes_read.sum"] + r["dram__bytes_write.sum"]
        tt += t; tb += b
        out.append(f"{r['short'][:52]:52s} {r['grid']:>14s} {t:8.2f} {r.get('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 0):6.1f} "
                   f"{r.get('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 0):6.1f} {b / 1e6:8.1f}")
        a = agg.setdefault(re.sub(r"<.*", "", r["short"]), [0, 0.0, 0.0, 0.0])
        a[0] += 1; a[1] += t; a[2] += b
        a[3] += t * r.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", 0)
    out.append(f"sum of durations {tt:.1f} us over {len(step)} launches; DRAM traffic {tb / 1e6:.1f} MB")
    out.append("")
    out.append("per kernel family: launches | us | % of the step's serialised kernel time | DRAM MB | time-weighted tensor pipe %")
    for k, (c, t, b, tw) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        out.append(f"  {k:28s} x{c:<3d} {t:8.1f} us {100 * t / tt:5.1f}%  {b / 1e6:8.1f} MB  {tw / max(t, 1e-9):5.1f}")
    tc = [r for r in step if r["short"].startswith(("gemm_tc_kernel", "enc_fwd_kernel"))]
    tct = sum(r["gpu__time_duration.sum"] for r in tc)
    tcb = sum(r["dram__bytes_read.sum"] + r["dram__bytes_write.sum"] for r in tc)
    tcw = sum(r["gpu__time_duration.sum"] * r.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", 0) for r in tc) / max(tct, 1e-9)
    big = [r for r in tc if r["short"].startswith("gemm_tc_kernel<256")]
    bigw = sum(r["gpu__time_duration.sum"] * r.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", 0) for r in big) / max(sum(r["gpu__time_duration.sum"] for r in big), 1e-9)
    out.append(f"tensor-core kernels (gemm_tc_kernel + enc_fwd_kernel): {len(tc)} launches, {tct:.1f} us = {100 * tct / tt:.1f}% of the step's kernel time, "
               f"time-weighted tensor pipe {tcw:.1f}% (the {len(big)} 256-wide GEMMs: {bigw:.1f}%), DRAM {tcb / 1e6:.1f} MB")
    os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)
    open(os.path.join(ROOT, "profiles", f"{tag}_ncu_step_summary.txt"), "w").write("\n".join(out) + "\n")
    json.dump({"cfg2_b4096": {"tensor_core_kernels_all_launches": {"launches": len(tc), "dram_bytes_per_step": tcb, "us_per_step": tct,
                                                                  "tensor_pipe_pct_time_weighted": tcw},
                              "whole_step": {"launches": len(step), "dram_bytes_per_step": tb, "serialised_kernel_us": tt}},
               "source": os.path.basename(fn), "how": "tools/profile_r2.sh -> tools/summarize_launches.py"},
              open(os.path.join(ROOT, "profiles", f"{tag}_traffic.json"), "w"), indent=1)
    print("\n".join(out[-12:]))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else "r2")
