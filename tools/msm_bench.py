#!/usr/bin/env python
"""G1/G2 MSM micro-benchmark (BASELINE.json configs[4]: sweep 2^16..2^28) on one GPU, inputs resident in HBM.

Bases are k_i*G for splitmix64-derived k_i (fixed-base kernel, so the result is checkable as (sum s_i k_i) G),
scalars uniform 256-bit (unreduced, top bits set, like depends/ffiasm/benchmark/multiexp_g1.cpp:12-32) or
reduced mod r (--scalars fr).  Prints one JSON line per size with points/s and the per-phase device times.
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import rapidsnark_old_b200 as b200
from rapidsnark_old_b200 import synth


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--log-n", type=int, nargs="+", default=[20])
    ap.add_argument("--g2", action="store_true")
    ap.add_argument("--c", type=int, default=0)
    ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--scalars", default="full", choices=["full", "fr"])
    ap.add_argument("--check", action="store_true")
    ap.add_argument("--acc-smem", type=int, default=-1)
    args = ap.parse_args()
    ctx = b200.Context(0)
    if args.c:
        ctx.set_msm_window(args.c)
    ctx.set_option("acc_smem", args.acc_smem)
    stream = torch.cuda.ExternalStream(ctx.stream())
    for log_n in args.log_n:
        n = 1 << log_n
        rng = np.random.default_rng(5)
        # distinct bases from a pool of 2^20 (larger n repeats the pool: same arithmetic cost, bounded setup)
        pool = min(n, 1 << 20)
        ks = rng.integers(0, 1 << 63, size=(pool, 4), dtype=np.uint64)
        ks[:, 3] &= (1 << 60) - 1
        g = synth.g2_gen_bytes() if args.g2 else synth.g1_gen_bytes()
        fb = ctx.fixed_base_g2 if args.g2 else ctx.fixed_base_g1
        pts = np.frombuffer(fb(g, ks.tobytes(), pool), dtype=np.uint8)
        psz = 128 if args.g2 else 64
        d_bases = torch.from_numpy(pts.copy()).cuda()
        if n > pool:
            d_bases = d_bases.repeat(n // pool)
        sc = rng.integers(0, 1 << 63, size=(n, 4), dtype=np.uint64) * 2 + 1
        if args.scalars == "fr":
            sc[:, 3] &= (1 << 61) - 1
        d_sc = torch.from_numpy(sc.view(np.uint8).reshape(-1).copy()).cuda()
        run = ctx.msm_g2_dev if args.g2 else ctx.msm_g1_dev
        out = run(d_bases.data_ptr(), d_sc.data_ptr(), n)      # warm-up + optional check
        if args.check and n <= (1 << 20):
            R = synth.R
            kk = [int.from_bytes(ks[i].tobytes(), "little") for i in range(pool)]
            ss = [int.from_bytes(sc[i].tobytes(), "little") for i in range(n)]
            tot = sum(a * b for a, b in zip(kk, ss)) % R
            if args.g2:
                assert b200.host_g2_to_affine(out) == b200.host_g2_to_affine(b200.host_g2_mul(g, tot.to_bytes(32, "little")))
            else:
                assert b200.host_g1_to_affine(out) == b200.host_g1_to_affine(b200.host_g1_mul(g, tot.to_bytes(32, "little")))
        for _ in range(2):
            run(d_bases.data_ptr(), d_sc.data_ptr(), n)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        phases = {}
        e0.record(stream)
        for _ in range(args.iters):
            run(d_bases.data_ptr(), d_sc.data_ptr(), n)
            for k, v in ctx.phase_ms().items():
                phases[k] = phases.get(k, 0.0) + v / args.iters
        e1.record(stream)
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / args.iters
        print(json.dumps({"group": "G2" if args.g2 else "G1", "log_n": log_n, "ms": round(ms, 4),
                          "points_per_s": round(n / ms * 1e3), "scalars": args.scalars, "c": args.c or "auto",
                          "phases_ms": {k: round(v, 4) for k, v in phases.items() if v > 0}}), flush=True)
        del d_bases, d_sc
    ctx.close()


if __name__ == "__main__":
    main()
