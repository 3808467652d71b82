#!/usr/bin/env python
"""Quick A/B of the experimental G2 accumulation (option acc_smem = 2) against the default on one GPU: same point?
how long does the accumulation phase take?  No torch, no oracle: starts in a second.  python tools/staged_check.py [log_n]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import rapidsnark_old_b200 as b
from rapidsnark_old_b200 import synth

log_n = int(sys.argv[1]) if len(sys.argv) > 1 else 18
c = b.Context(0)
n = 1 << log_n
rng = np.random.default_rng(1)
ks = rng.integers(0, 1 << 63, size=(n, 4), dtype=np.uint64)
ks[:, 3] &= (1 << 60) - 1
pts = c.fixed_base_g2(synth.g2_gen_bytes(), ks.tobytes(), n)
sc = rng.integers(0, 1 << 63, size=(n, 4), dtype=np.uint64)
sc[:, 3] &= (1 << 61) - 1
scb = sc.tobytes()
res = {}
for v in (0, 2, 0, 2, 1):
    c.set_option("acc_smem", v)
    t = time.time()
    r = c.msm_g2(pts, scb, n)
    aff = b.host_g2_to_affine(r).hex()
    res.setdefault(v, aff)
    print("acc_smem=%d" % v, aff[:24], "accumulate_g2 %.3f ms" % c.phase_ms().get("msm_accumulate_g2", 0.0),
          "call %.1f ms" % ((time.time() - t) * 1e3), flush=True)
print("SAME POINT" if len(set(res.values())) == 1 else "MISMATCH", flush=True)
