#!/usr/bin/env python
"""Proofs per second against ONE resident zkey (the reference's server use, src/fullprover.cpp:84-99).

  python tools/throughput_bench.py [--log-n 20] [--contexts 1 2 3] [--seconds 3] [--gpus N]

Every context is a b200_ctx with a view (b200_zkey_share) of the same device-resident tables and is driven by its own
host thread calling b200_groth16_prove in a loop with the witness in pinned host memory (H2D inside every proof).
With K > 1 contexts the GPU overlaps one proof's tail - last bucket reduction, read-back, host finalisation - and the
next proof's start (witness upload, digit sort).  --gpus N: N independent replicas of that (one zkey per GPU, proofs
are independent objects: "weak" scaling, no collective).  Every proof is checked against the first one (same r, s).
Prints one JSON line per configuration.
"""
import argparse
import ctypes
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--log-n", type=int, default=20)
    ap.add_argument("--contexts", type=int, nargs="*", default=[1, 2, 3])
    ap.add_argument("--seconds", type=float, default=3.0)
    ap.add_argument("--gpus", type=int, default=1)
    args = ap.parse_args()
    import torch
    import bench
    import rapidsnark_old_b200 as b200

    if not torch.cuda.is_available():
        raise SystemExit("throughput_bench: no CUDA device")
    ctx0 = b200.Context(0)
    s = bench.build_inputs(args.log_n, 2, *bench.gpu_point_makers(ctx0))
    p, vk = s.points, s.vk
    coefs = s.coefs_section()
    wt = s.wtns_bytes()
    wt_host = torch.empty(len(wt), dtype=torch.uint8).pin_memory()
    wt_host.copy_(torch.frombuffer(bytearray(wt), dtype=torch.uint8))
    r32, s32 = bench.blinding_factors()
    kmax = max(args.contexts)
    # per GPU: one owner context with the tables, kmax - 1 views
    ctxs, zks = [], []
    for g in range(args.gpus):
        owner = ctx0 if g == 0 else b200.Context(g)
        zk = owner.zkey_upload(s.n_vars, s.n_public, s.n, s.n_coefs, coefs, p["A"], p["B1"], p["B2"], p["C"], p["H"])
        cs, zs = [owner], [zk]
        for _ in range(kmax - 1):
            c = b200.Context(g)
            cs.append(c)
            zs.append(c.zkey_share(zk))
        ctxs.append(cs)
        zks.append(zs)
    _, want = zks[0][0].prove(wt_host.data_ptr(), vk, r32, s32)
    bench.check_known_dlogs(b200, s, None, want, r32, s32)
    for k in args.contexts:
        for g in range(args.gpus):                      # warm-up: every context once (workspaces, twiddles)
            for z in zks[g][:k]:
                assert z.prove(wt_host.data_ptr(), vk, r32, s32)[1] == want
        counts = [[0] * k for _ in range(args.gpus)]
        bad = []
        start = threading.Barrier(args.gpus * k + 1)
        deadline = [0.0]

        def worker(g, i):
            z = zks[g][i]
            start.wait()
            while time.perf_counter() < deadline[0]:
                if z.prove(wt_host.data_ptr(), vk, r32, s32)[1] != want:
                    bad.append((g, i))
                counts[g][i] += 1

        th = [threading.Thread(target=worker, args=(g, i)) for g in range(args.gpus) for i in range(k)]
        for t in th:
            t.start()
        t0 = time.perf_counter()
        deadline[0] = t0 + args.seconds
        start.wait()
        for t in th:
            t.join()
        dt = time.perf_counter() - t0
        total = sum(sum(c) for c in counts)
        assert not bad, "wrong proofs from %s" % bad
        print(json.dumps({"metric": "groth16_proofs_per_s_2^%d_constraints" % args.log_n, "value": round(total / dt, 2),
                          "unit": "proofs/s", "n_gpus": args.gpus, "contexts_per_gpu": k, "proofs": total,
                          "seconds": round(dt, 3), "ms_per_proof_per_gpu": round(dt * 1e3 * args.gpus / total, 3),
                          "scaling": "weak (independent replicas)" if args.gpus > 1 else "n/a",
                          "h2d_bytes_per_proof": len(wt), "checked": "every proof equals the first (same r, s), "
                          "which is verified against its known discrete logs"}), flush=True)
    for g in range(args.gpus):
        for z in reversed(zks[g]):
            z.free()
        for c in ctxs[g]:
            c.close()
    return 0


if __name__ == "__main__":
    sys.exit(main())
