#!/usr/bin/env python
"""Tuning harness: one synthetic 2^k circuit, several library configurations, per-phase device times.
  python tools/prove_bench.py --log-n 20 --configs precomp=0 precomp_c=14 precomp_c=16,acc_smem=0
"""
import argparse
import hashlib
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import rapidsnark_old_b200 as b200
import bench


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--log-n", type=int, default=20)
    ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--configs", nargs="*", default=[""])
    args = ap.parse_args()
    ctx = b200.Context(0)
    s = bench.build_inputs(args.log_n, 2, *bench.gpu_point_makers(ctx))
    p = s.points
    wt = s.wtns_bytes()
    wt_dev = torch.frombuffer(bytearray(wt), dtype=torch.uint8).cuda()
    stream = torch.cuda.ExternalStream(ctx.stream())
    coefs = s.coefs_section()
    first = None
    for cfg in args.configs:
        opts = dict(kv.split("=") for kv in cfg.split(",") if kv)
        for k in ("precomp", "precomp_c", "acc_smem", "msm_window", "target_tasks_log2", "h_streams", "g2_minb", "warm_max", "reduce_l", "reduce_l_g2", "tree_threads", "lockstep_g1", "lockstep_g2", "fuse_g1", "h_early", "reduce_l_tail", "plane_items"):
            ctx.set_option(k, int(opts.get(k, -1 if k in ("precomp", "acc_smem", "fuse_g1", "h_early", "lockstep_g1") else 0)))
        shards = int(opts.get("shards", 1))     # this GPU plays rank 0 of `shards` (per-rank time of an N-GPU run)
        t0 = time.time()
        zk = ctx.zkey_upload(s.n_vars, s.n_public, s.n, s.n_coefs, coefs, p["A"], p["B1"], p["B2"], p["C"], p["H"],
                             0, shards)
        up = time.time() - t0
        for _ in range(2):
            out = zk.prove_msms_dev(wt_dev.data_ptr())
        torch.cuda.synchronize()
        # same five points as the first configuration (canonical affine values)?
        aff = [b200.host_g1_to_affine(out[0:128]), b200.host_g1_to_affine(out[128:256]), b200.host_g1_to_affine(out[256:384]),
               b200.host_g2_to_affine(out[384:640]), b200.host_g1_to_affine(out[640:768])]
        if first is None:
            first = aff
        same = aff == first
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ph = {}
        e0.record(stream)
        for _ in range(args.iters):
            zk.prove_msms_dev(wt_dev.data_ptr())
            for k, v in ctx.phase_ms().items():
                ph[k] = ph.get(k, 0.0) + v / args.iters
        e1.record(stream)
        torch.cuda.synchronize()
        print(json.dumps({"config": cfg or "default", "ms": round(e0.elapsed_time(e1) / args.iters, 3),
                          "upload_s": round(up, 2), "same_points_as_first": same, "points_sha": hashlib.sha256(b"".join(aff)).hexdigest()[:16], "phases": {k: round(v, 3) for k, v in ph.items() if v > 0},
                          "free_gb": round(torch.cuda.mem_get_info()[0] / 2**30, 1)}), flush=True)
        zk.free()
    ctx.close()


if __name__ == "__main__":
    main()
