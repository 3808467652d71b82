#!/usr/bin/env python
"""Write the synthetic circuit of SURVEY.md Appendix C (zkey tables, coefficient section, witness, key) as raw files.

  python tools/make_inputs.py --log-n 20 --seed 2 --out DIR

Used by `bench.py --impl reference` so that the process which TIMES the reference's CPU prover never loads
libb200snark.so: the tables are produced here, in a separate process (on the GPU by the library's fixed-base
routine when a device is present, by the CPU oracle otherwise - setup only, never timed), and read back from disk.
Files: A.bin B1.bin B2.bin C.bin H.bin (affine Montgomery points), coefs.bin (zkey section 4), wtns.bin (normal-form
scalars), meta.json (sizes, verification-key points as hex, known discrete logs).
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--log-n", type=int, default=20)
    ap.add_argument("--seed", type=int, default=2)
    ap.add_argument("--out", required=True)
    args = ap.parse_args()
    import rapidsnark_old_b200 as b200
    from rapidsnark_old_b200 import synth
    t0 = time.time()
    try:
        ctx = b200.Context(0)
        g1, g2 = synth.g1_gen_bytes(), synth.g2_gen_bytes()
        makers = (lambda ks: ctx.fixed_base_g1(g1, synth.le32_many(ks), synth.count32(ks)),
                  lambda ks: ctx.fixed_base_g2(g2, synth.le32_many(ks), synth.count32(ks)))
        how = "gpu fixed-base"
    except Exception as e:
        import synth_util
        ctx, makers, how = None, synth_util.oracle_point_makers(), "cpu oracle (%s)" % e
    # without a GPU the oracle makes the tables from Python integers: the plain Synth (same values, slower)
    s = synth.FastSynth(args.log_n, args.seed) if ctx else synth.Synth(args.log_n, args.seed)
    s.build_points(*makers)
    if ctx:
        ctx.close()
    os.makedirs(args.out, exist_ok=True)
    for k in ("A", "B1", "B2", "C", "H"):
        with open(os.path.join(args.out, k + ".bin"), "wb") as f:
            f.write(s.points[k])
    with open(os.path.join(args.out, "coefs.bin"), "wb") as f:
        f.write(s.coefs_section())
    with open(os.path.join(args.out, "wtns.bin"), "wb") as f:
        f.write(s.wtns_bytes())
    meta = {"log_n": args.log_n, "seed": args.seed, "n": s.n, "n_vars": s.n_vars, "n_public": s.n_public,
            "n_coefs": s.n_coefs, "vk": {k: v.hex() for k, v in s.vk.items()},
            "dlog_a": str(s.dlog_a), "dlog_b": str(s.dlog_b), "dlog_pub": str(s.dlog_pub),
            "alpha": str(s.alpha), "beta": str(s.beta), "delta": str(s.delta), "tables_by": how}
    with open(os.path.join(args.out, "meta.json"), "w") as f:
        json.dump(meta, f)
    print("[make_inputs] 2^%d circuit written to %s in %.1fs (%s)" % (args.log_n, args.out, time.time() - t0, how),
          file=sys.stderr)
    return 0


if __name__ == "__main__":
    sys.exit(main())
