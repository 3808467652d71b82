#!/usr/bin/env python
"""Turns the ncu outputs brought back in gpurun_out/ into the tracked summaries under profiles/.
  python tools/summarize_profiles.py r01
reads  gpurun_out/<tag>_launches.csv   (ncu --metrics gpu__time_duration.sum launch list of `python bench.py`)
       gpurun_out/<tag>_accumulate.ncu-rep   (ncu --set full of k_msm_accumulate*)
writes profiles/<tag>_launches.csv (copy), profiles/<tag>_launch_summary.md, profiles/<tag>_accumulate_metrics.json/.md
"""
import collections
import csv
import json
import os
import re
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
src, dst = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
os.makedirs(dst, exist_ok=True)


def short(name):
    name = re.sub(r"\(.*", "", name)
    name = name.replace("void ", "").replace("b200::", "")
    name = name.replace("Fq2T<Fp<FqParams>>", "Fq2").replace("Fp<FqParams>", "Fq").replace("Fp<FrParams>", "Fr")
    return name


lc = os.path.join(src, tag + "_launches.csv")
if os.path.exists(lc):
    shutil.copy(lc, os.path.join(dst, tag + "_launches.csv"))
    rows = [r for r in csv.reader(open(lc)) if len(r) > 8]
    hdr = rows[0]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    data = []
    for r in rows[1:]:
        v = float(r[vi].replace(",", ""))
        ms = v / 1e6 if r[ui] == "ns" else v / 1e3 if r[ui] in ("us", "usecond") else v
        data.append((short(r[ki]), ms))
    # one proof = the last `period` launches, period = distance between consecutive k_build_abc launches
    # (the witness sort and the G2 accumulation are enqueued before the H pipeline, so a proof does not start there)
    starts = [i for i, d in enumerate(data) if d[0].startswith("k_build_abc")]
    period = starts[-1] - starts[-2] if len(starts) > 1 else len(data)
    last = data[len(data) - period:]
    agg = collections.OrderedDict()
    for n, ms in last:
        a = agg.setdefault(n, [0, 0.0])
        a[0] += 1
        a[1] += ms
    tot = sum(ms for _, ms in last)
    with open(os.path.join(dst, tag + "_launch_summary.md"), "w") as f:
        f.write("# %s: kernels of ONE proof (2^20 constraints, 1 x B200), ncu launch list of `python bench.py`\n\n" % tag)
        f.write("Command: `ncu --metrics gpu__time_duration.sum --clock-control none --csv python bench.py --steps 2 --warmup 3`.\n"
                "Per-launch times are serialised and cold-cache (compare SHARES, not absolutes; the live run overlaps the\n"
                "side-stream kernels with the accumulations).  Full list: `%s_launches.csv`.\n\n" % tag)
        f.write("| kernel | launches | total ms | share |\n|---|---:|---:|---:|\n")
        for n, (cnt, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write("| `%s` | %d | %.3f | %.1f%% |\n" % (n, cnt, ms, 100 * ms / tot))
        f.write("| **total** | %d | %.3f | 100%% |\n" % (len(last), tot))
    print("launch summary:", len(last), "launches,", round(tot, 3), "ms")

rep = os.path.join(src, tag + "_accumulate.ncu-rep")
raw = os.path.join(src, tag + "_accumulate_raw.csv")      # `ncu -i <rep> --page raw --csv`, made on the GPU box
if os.path.exists(rep) or os.path.exists(raw):
    if os.path.exists(raw):
        out = open(raw).read()
    else:
        out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread",
            "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers",
            "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed",
            "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
            "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "smsp__inst_executed.sum",
            "dram__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
            "sm__throughput.avg.pct_of_peak_sustained_elapsed",
            "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio"]
    kernels = []
    for r in rows[2:]:
        d = {"kernel": short(r[hdr.index("Kernel Name")])}
        for w in want:
            if w in hdr:
                i = hdr.index(w)
                try:
                    d[w] = float(r[i].replace(",", ""))
                except ValueError:
                    d[w] = r[i]
                d[w + "__unit"] = units[i]
        kernels.append(d)
    json.dump(kernels, open(os.path.join(dst, tag + "_accumulate_metrics.json"), "w"), indent=1)
    with open(os.path.join(dst, tag + "_accumulate_metrics.md"), "w") as f:
        f.write("# %s: `ncu --set full --clock-control none --import-source on -k regex:k_msm_accumulate` on `python bench.py`\n\n" % tag)
        for d in kernels:
            f.write("## %s\n\n| metric | value | unit |\n|---|---:|---|\n" % d["kernel"])
            for w in want:
                if w in d:
                    f.write("| %s | %s | %s |\n" % (w, d[w], d[w + "__unit"]))
            f.write("\n")
    print("accumulate metrics for", len(kernels), "launches")
    # DRAM traffic per launch of the G1 accumulation kernel, tagged with the commit and configuration of the capture:
    # bench.py prints it as roofline.traffic only for a run of the same configuration
    unit = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
    g1 = [k for k in kernels if k["kernel"].startswith("k_msm_accumulate") and "Fq2" not in k["kernel"] and "<Fq" in k["kernel"]]
    if g1:
        tot = [k["dram__bytes_read.sum"] * unit[k["dram__bytes_read.sum__unit"]] +
               k["dram__bytes_write.sum"] * unit[k["dram__bytes_write.sum__unit"]] for k in g1]
        # the witness MSMs' launch (three MSMs) and the H launch differ in size: report the mean per launch, as
        # bench.py's algorithmic bytes per launch are a mean as well
        commit = subprocess.run(["git", "rev-parse", "--short", "HEAD"], capture_output=True, text=True, cwd=ROOT).stdout.strip()
        meta = {"kernel": g1[0]["kernel"], "launches_captured": len(g1), "dram_bytes_per_launch": round(sum(tot) / len(tot)),
                "dram_bytes_each": [round(t) for t in tot], "log_n": int(os.environ.get("B200_CAPTURE_LOG_N", "20")),
                "n_gpus": int(os.environ.get("B200_CAPTURE_GPUS", "1")), "commit": commit,
                "command": "ncu --set full --clock-control none --import-source on -k regex:k_msm_accumulate python bench.py --steps 1 --warmup 3"}
        json.dump(meta, open(os.path.join(dst, tag + "_accumulate_traffic.json"), "w"), indent=1)
        print("traffic:", meta["dram_bytes_per_launch"], "bytes per launch over", len(g1), "launches")
