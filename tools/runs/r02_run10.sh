#!/bin/bash
# round 2, GPU call 10: NTT changes (precombined twist table, TMA passes) - parity and A/B
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_benchsize.py -m gpu -x -q -k "ntt or h_scalars or tma" 2>&1 | tail -6 )
python - <<'PY'
import sys, time, ctypes
sys.path.insert(0, ".")
import numpy as np, torch
import rapidsnark_old_b200 as b200
ctx = b200.Context(0)
for log_n in (20, 24):
    n = 1 << log_n
    data = torch.from_numpy(np.random.default_rng(1).integers(0, 1 << 61, size=(n, 4), dtype=np.uint64).view(np.int64)).cuda()
    for tma in (0, 1):
        ctx.set_option("ntt_tma", tma)
        for inv in (False, True):
            for _ in range(3): ctx.ntt_dev(data.data_ptr(), n, inv)
            torch.cuda.synchronize(); t = time.perf_counter()
            for _ in range(10): ctx.ntt_dev(data.data_ptr(), n, inv)
            torch.cuda.synchronize()
            print("ntt 2^%d tma=%d inverse=%d: %.3f ms" % (log_n, tma, inv, (time.perf_counter() - t) * 100), ctx.phase_ms().get("ntt_h"))
PY
for t in 0 1; do
  timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --opt ntt_tma=$t > gpurun_out/r02_tma$t.json 2> gpurun_out/r02_tma$t.log
  python - gpurun_out/r02_tma$t.json <<'PY'
import json,sys
d=json.loads([l for l in open(sys.argv[1]) if l.startswith("{")][-1])
print(sys.argv[1], d["value"], d["e2e"]["value"], d["timeline_ms"]["ntt_h"], d["timeline_ms"]["_span"])
PY
done
