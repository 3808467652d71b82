#!/bin/bash
# round 2, GPU call 15: emulated rank 0 of 2 and of 4 on one GPU under the two-lane bucket reduction - is one fused
# launch for all four G1 accumulations still right, and which segment length
mkdir -p gpurun_out
emu() { # shards tag opts...
  n=$1; tag=$2; shift 2
  timeout 120 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --emulate-shards $n --emulate-rank 0 --emulate-poly-mask -2 --opt "$@" > gpurun_out/r02_run15_emu${n}_$tag.json 2> gpurun_out/r02_run15_emu${n}_$tag.log
  python - gpurun_out/r02_run15_emu${n}_$tag.json <<'PY'
import json,sys
try:
    d=json.loads([l for l in open(sys.argv[1]) if l.startswith("{")][-1])
    t=d.get("timeline_ms", {})
    print(sys.argv[1], "value", d["value"], "span", t.get("_span"), "acc_g1 end", t.get("msm_accumulate_g1", {}).get("end"), "reduce end", t.get("msm_reduce", {}).get("end"))
except Exception as e:
    print(sys.argv[1], "failed", e)
PY
}
emu 2 default timeline=0
emu 2 fuse3 fuse_g1=3
emu 2 l16 reduce_l=16
emu 2 l32 reduce_l=32
emu 4 default timeline=0
emu 4 fuse3 fuse_g1=3
emu 4 l16 reduce_l=16
emu 2 fuse3_l16 fuse_g1=3 reduce_l=16
