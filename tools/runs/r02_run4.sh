#!/bin/bash
# round 2, GPU call 4: fused G1 accumulation launch - parity, then A/B at N = 1 and on an emulated rank 0 of 8
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_golden.py -m gpu -x -q 2>&1 | tail -4 )
show() { python - "$1" <<'PY'
import json,sys
try:
    d=json.loads([l for l in open(sys.argv[1]) if l.startswith("{")][-1])
    t=d["timeline_ms"]
    print(sys.argv[1].split("/")[-1], "value", d["value"], "e2e", d["e2e"]["value"], "circom", (d.get("circom_like_witness") or {}).get("value"),
          "| acc_g2", t.get("msm_accumulate_g2"), "acc_g1", t.get("msm_accumulate_g1"), "reduce", t.get("msm_reduce"), "ntt", t.get("ntt_h"), "span", t["_span"])
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
}
for f in -1 3 0; do
  timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --opt fuse_g1=$f > gpurun_out/r02_fuse${f}_n1.json 2> gpurun_out/r02_fuse${f}_n1.log
  show gpurun_out/r02_fuse${f}_n1.json
done
for f in -1 0; do
  timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --emulate-shards 8 --emulate-poly-mask 1 --opt fuse_g1=$f > gpurun_out/r02_fuse${f}_emu8.json 2> gpurun_out/r02_fuse${f}_emu8.log
  show gpurun_out/r02_fuse${f}_emu8.json
  timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --emulate-shards 4 --emulate-poly-mask 1 --opt fuse_g1=$f > gpurun_out/r02_fuse${f}_emu4.json 2> gpurun_out/r02_fuse${f}_emu4.log
  show gpurun_out/r02_fuse${f}_emu4.json
done
tail -2 gpurun_out/r02_fuse-1_emu8.log
