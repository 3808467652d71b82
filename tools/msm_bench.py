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
    ap.add_argument("--scalars", default="full", choices=["full", "fr", "circom"],
                    help="full: uniform 256-bit; fr: uniform below r; circom: 70 %% of the scalars in {0,1}, 20 %% < 2^32, "
                         "10 %% uniform below r (SURVEY.md 8d config 2, the shape of real circom witnesses)")
    ap.add_argument("--check", action="store_true")
    ap.add_argument("--acc-smem", type=int, default=-1)
    ap.add_argument("--cpu-max-log-n", type=int, default=0,
                    help="also time the reference's CPU multiexp (oracle/_ref, all host cores) up to this size (1 GPU runs only)")
    args = ap.parse_args()
    # under torchrun (WORLD_SIZE > 1): the points are sharded by range over the ranks, each rank runs the MSM of its
    # slice, the 128 / 256-byte partial sums are all_gathered (NCCL) and folded on the host; time = max over ranks
    import torch.distributed as dist
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = b200.Context(local)
    if args.c:
        ctx.set_msm_window(args.c)
    ctx.set_option("acc_smem", args.acc_smem)
    stream = torch.cuda.ExternalStream(ctx.stream(), device=torch.device("cuda", local))
    for log_n in args.log_n:
        n_total = 1 << log_n
        n = n_total * (rank + 1) // world - n_total * rank // world      # this rank's slice
        rng = np.random.default_rng(5 + rank)
        # distinct bases from a pool of 2^20 (larger n repeats the pool: same arithmetic cost, bounded setup)
        pool = min(n, 1 << 20)
        psz_out = 256 if args.g2 else 128
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
        if args.scalars in ("fr", "circom"):
            sc[:, 3] &= (1 << 61) - 1
        if args.scalars == "circom":
            kind = rng.random(n)
            small = kind < 0.7
            mid = (kind >= 0.7) & (kind < 0.9)
            sc[small] = 0
            sc[small, 0] = rng.integers(0, 2, size=int(small.sum()), dtype=np.uint64)
            sc[mid] = 0
            sc[mid, 0] = rng.integers(0, 1 << 32, size=int(mid.sum()), dtype=np.uint64)
        d_sc = torch.from_numpy(sc.view(np.uint8).reshape(-1).copy()).cuda()
        run = ctx.msm_g2_dev if args.g2 else ctx.msm_g1_dev
        out = run(d_bases.data_ptr(), d_sc.data_ptr(), n)      # warm-up + optional check
        if args.check and n <= (1 << 20) and world == 1:
            R = synth.R
            kk = [int.from_bytes(ks[i].tobytes(), "little") for i in range(pool)]
            ss = [int.from_bytes(sc[i].tobytes(), "little") for i in range(n)]
            tot = sum(a * b for a, b in zip(kk, ss)) % R
            if args.g2:
                assert b200.host_g2_to_affine(out) == b200.host_g2_to_affine(b200.host_g2_mul(g, tot.to_bytes(32, "little")))
            else:
                assert b200.host_g1_to_affine(out) == b200.host_g1_to_affine(b200.host_g1_mul(g, tot.to_bytes(32, "little")))
        def step():
            part = run(d_bases.data_ptr(), d_sc.data_ptr(), n)
            if world > 1:
                mine = torch.frombuffer(bytearray(part), dtype=torch.uint8).cuda()
                outs = [torch.empty_like(mine) for _ in range(world)]
                dist.all_gather(outs, mine)
                acc = bytes(outs[0].cpu().numpy())
                for o in outs[1:]:
                    acc = (b200.host_g2_add if args.g2 else b200.host_g1_add)(acc, bytes(o.cpu().numpy()))
                return acc
            return part
        for _ in range(2):
            step()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        phases = {}
        e0.record(stream)
        for _ in range(args.iters):
            step()
            for k, v in ctx.phase_ms().items():
                phases[k] = phases.get(k, 0.0) + v / args.iters
        e1.record(stream)
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / args.iters
        if world > 1:
            t = torch.tensor([ms], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        cpu = None
        if world == 1 and log_n <= args.cpu_max_log_n:
            sys.path.insert(0, os.path.join(ROOT, "tests"))
            import oracle_lib
            o = oracle_lib.ref() or oracle_lib.port()
            hb, hs = bytes(d_bases.cpu().numpy()), bytes(d_sc.cpu().numpy())
            msm, add = (o.g2_msm, o.g2_add) if args.g2 else (o.g1_msm, o.g1_add)
            t0 = time.perf_counter()
            if n >= (1 << 27):
                # the reference computes scalarIdx * scalarSize in 32 bits (multiexp.cpp:30): from 2^27 points of 32-byte
                # scalars on it reads the wrong scalars, so the call is split in halves and the two results added (flagged)
                h = n // 2
                ref_out = add(msm(hb[:h * psz], hs[:h * 32], h), msm(hb[h * psz:], hs[h * 32:], n - h))
            else:
                ref_out = msm(hb, hs, n)
            cpu_ms = (time.perf_counter() - t0) * 1e3
            same = (o.g2_to_affine if args.g2 else o.g1_to_affine)(ref_out) == (o.g2_to_affine if args.g2 else o.g1_to_affine)(out)
            cpu = {"ms": round(cpu_ms, 1), "points_per_s": round(n / cpu_ms * 1e3), "cores": o.threads(), "kind": o.kind,
                   "same_point_as_gpu": bool(same), "split_in_two_calls": n >= (1 << 27)}
        if rank == 0:
            print(json.dumps({"group": "G2" if args.g2 else "G1", "log_n": log_n, "n_gpus": world, "ms": round(ms, 4), "cpu": cpu,
                          "points_per_s": round(n_total / ms * 1e3), "scalars": args.scalars, "c": args.c or "auto",
                          "phases_ms": {k: round(v, 4) for k, v in phases.items() if v > 0}}), flush=True)
        del d_bases, d_sc
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
