#!/usr/bin/env python
"""Build record of libb200snark.so for profiles/: per kernel the ptxas figures (registers, spills, stack, shared
memory - from build/csrc/*.ptxas.log, written by the Makefile's `-Xptxas -v`), the SASS size, and for the two hot
accumulation loops the opcode mix of ONE loop iteration (= one mixed addition), so that claims like "1147 IMAD.WIDE
per G1 addition" can be checked.  No GPU needed.   python tools/build_record.py > profiles/rNN_build_record.md"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OBJ = os.path.join(ROOT, "build", "csrc")


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.split("\n")
    return dict(zip(names, out))


def ptxas_table(log):
    rows, cur = [], None
    for line in open(log):
        m = re.search(r"Compiling entry function '(\S+)'", line)
        if m:
            cur = {"name": m.group(1)}
            rows.append(cur)
            continue
        if cur is None:
            continue
        m = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", line)
        if m and "stack" not in cur:
            cur["stack"], cur["st"], cur["ld"] = m.groups()
        m = re.search(r"Used (\d+) registers", line)
        if m:
            cur["regs"] = m.group(1)
            s = re.search(r"(\d+) bytes smem", line)
            cur["smem"] = s.group(1) if s else "0"
            cur = None
    return rows


def sass_functions(obj):
    txt = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
    fns = {}
    for fn in re.split(r"\n\s*Function : ", txt)[1:]:
        name = fn.split("\n")[0].strip()
        ins = [(int(m.group(1), 16), m.group(2)) for m in re.finditer(r"^\s+/\*([0-9a-f]+)\*/\s+(.*?);", fn, re.M)]
        fns[name] = ins
    return fns


def loop_mix(ins):
    """opcode mix of the longest backward-branch loop of a kernel"""
    best = None
    for a, t in ins:
        m = re.search(r"BRA\s+(0x[0-9a-f]+)", t)
        if m and int(m.group(1), 16) < a and (best is None or a - int(m.group(1), 16) > best[1] - best[0]):
            best = (int(m.group(1), 16), a)
    if not best:
        return None
    c = collections.Counter()
    for a, t in ins:
        if best[0] <= a <= best[1]:
            op = re.sub(r"^@!?U?P\d+\s+", "", t).strip().split()[0]
            k = op.split(".")[0]
            if k == "IMAD":
                k = "IMAD.WIDE" if ".WIDE" in op else "IMAD.HI" if ".HI" in op else "IMAD.MOV" if ".MOV" in op else \
                    "IMAD.X" if (".X" in op or ".IADD" in op) else "IMAD"
            c[k] += 1
    return (best[1] - best[0]) // 16 + 1, c


def short(n):
    n = re.sub(r"b200::", "", n)
    n = n.replace("Fq2T<Fp<FqParams> >", "Fq2").replace("Fp<FqParams>", "Fq").replace("Fp<FrParams>", "Fr")
    return re.sub(r"\(.*", "", n).replace("void ", "")


def main():
    commit = subprocess.run(["git", "-C", ROOT, "rev-parse", "--short", "HEAD"], capture_output=True, text=True).stdout.strip()
    print("# Build record: libb200snark.so kernels (nvcc 12.9, -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo)\n")
    print("Made by `python tools/build_record.py` at commit `%s` (+ working tree) from `build/csrc/*.ptxas.log` and" % commit)
    print("`cuobjdump -sass build/csrc/*.o`.  SASS instructions are 16 bytes each; the SM's instruction cache is 32 KB.\n")
    for tu in ("msm_g1", "msm_g2", "ntt", "prove", "capi"):
        log, obj = os.path.join(OBJ, tu + ".ptxas.log"), os.path.join(OBJ, tu + ".o")
        if not os.path.exists(log):
            continue
        rows = ptxas_table(log)
        if not rows:
            continue
        fns = sass_functions(obj)
        names = demangle([r["name"] for r in rows])
        print("## %s.cu\n\n| kernel | registers | spill st/ld (B) | stack (B) | static smem (B) | SASS instr | KB |" % tu)
        print("|---|---:|---:|---:|---:|---:|---:|")
        for r in rows:
            n = len(fns.get(r["name"], []))
            print("| `%s` | %s | %s / %s | %s | %s | %d | %.1f |" % (short(names[r["name"]]), r.get("regs", "?"), r.get("st", "?"),
                                                                  r.get("ld", "?"), r.get("stack", "?"), r.get("smem", "0"), n, n * 16 / 1024))
        print()
        hot = [r for r in rows if re.search(r"k_msm_accumulate", r["name"])]
        if hot:
            print("Loop of the accumulation kernels (one iteration = one mixed addition; the longest backward branch):\n")
            print("| kernel | loop instr | loop KB | IMAD.WIDE | IMAD.HI | IMAD(.X/.MOV) | IADD3 | LDG | LDL+STL | BAR | other |")
            print("|---|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|")
            for r in hot:
                lm = loop_mix(fns.get(r["name"], []))
                if not lm:
                    continue
                n, c = lm
                imad = c["IMAD"] + c["IMAD.X"] + c["IMAD.MOV"]
                lmem = sum(v for k, v in c.items() if k in ("LDL", "STL"))
                known = c["IMAD.WIDE"] + c["IMAD.HI"] + imad + c["IADD3"] + c["LDG"] + lmem + c["BAR"]
                print("| `%s` | %d | %.1f | %d | %d | %d | %d | %d | %d | %d | %d |" % (short(names[r["name"]]), n, n * 16 / 1024, c["IMAD.WIDE"],
                                                                                    c["IMAD.HI"], imad, c["IADD3"], c["LDG"], lmem, c["BAR"], n - known))
            print()


if __name__ == "__main__":
    main()
