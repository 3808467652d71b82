#!/usr/bin/env python
"""Warp-stall sampling of an ncu report (--set full --import-source on) broken down by stall reason and by SASS opcode,
per kernel launch.   python tools/ncu_stalls.py gpurun_out/r01_accumulate.ncu-rep [out.md]
(reads `ncu -i <rep> --page source --csv`; no GPU needed)"""
import collections
import csv
import subprocess
import sys

STALLS = ["stall_wait", "stall_math", "stall_no_inst", "stall_selected", "stall_not_selected", "stall_dispatch",
          "stall_long_sb", "stall_short_sb", "stall_lg", "stall_mio", "stall_branch_resolving", "stall_barrier",
          "stall_drain", "stall_membar", "stall_misc", "stall_sleep", "stall_tex"]


def main():
    rep = sys.argv[1]
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-kernel-base", "function"],
                         capture_output=True, text=True).stdout
    kernels, cur = [], None
    for r in csv.reader(out.splitlines()):
        if r and r[0] == "Kernel Name":
            cur = {"name": r[1], "hdr": None, "rows": []}
            kernels.append(cur)
        elif cur is not None and cur["hdr"] is None:
            cur["hdr"] = r
        elif cur is not None:
            cur["rows"].append(r)
    lines = ["# Warp-stall samples by reason and by opcode (`%s`)\n" % rep,
             "`stall_selected` = the sample found the warp issuing; `stall_math` = the pipe it needs is busy (the good "
             "kind of stall for a pipe-bound kernel); `stall_wait` = fixed-latency dependency on the warp's own previous "
             "instruction; `stall_no_inst` = waiting for instruction fetch.\n"]
    seen = set()
    for k in kernels:
        h = k["hdr"]
        idx = {n: i for i, n in enumerate(h)}
        tot, byop, n_instr = collections.Counter(), collections.defaultdict(collections.Counter), 0
        for r in k["rows"]:
            if len(r) <= idx["stall_wait"]:
                continue
            toks = r[idx["Source"]].strip().split()
            if not toks:
                continue
            op = toks[1] if toks[0].startswith("@") and len(toks) > 1 else toks[0]
            key = op.split(".")[0] + (".WIDE" if ".WIDE" in op else "") + (".HI" if ".HI" in op else "")
            n_instr += 1
            for s in STALLS:
                if s in idx:
                    v = int(r[idx[s]] or 0)
                    tot[s] += v
                    byop[key][s] += v
        total = sum(tot.values())
        sig = (k["name"], n_instr)
        if sig in seen or total == 0:
            continue                      # repeated launches of the same kernel with the same profile
        seen.add(sig)
        lines.append("## %s - %d SASS instructions, %d samples\n" % (k["name"], n_instr, total))
        lines.append("| stall reason | share |\n|---|---:|")
        for s, v in tot.most_common():
            if v * 200 > total:
                lines.append("| %s | %.1f %% |" % (s, 100 * v / total))
        lines.append("\n| opcode | share of samples | main reasons |\n|---|---:|---|")
        for op, c in sorted(byop.items(), key=lambda kv: -sum(kv[1].values()))[:8]:
            t = sum(c.values())
            why = ", ".join("%s %d %%" % (s.replace("stall_", ""), round(100 * v / t)) for s, v in c.most_common(4) if v * 20 > t)
            lines.append("| %s | %.1f %% | %s |" % (op, 100 * t / total, why))
        lines.append("")
    text = "\n".join(lines) + "\n"
    if len(sys.argv) > 2:
        open(sys.argv[2], "w").write(text)
    print(text)


if __name__ == "__main__":
    main()
