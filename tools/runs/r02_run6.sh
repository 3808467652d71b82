#!/bin/bash
# round 2, GPU call 6: reduce-segment length / tree size sweep on an emulated rank 0 of 8 (and 4)
mkdir -p gpurun_out
show() { python - "$1" "$2" <<'PY'
import json,sys
try:
    d=json.loads([l for l in open(sys.argv[1]) if l.startswith("{")][-1])
    t=d["timeline_ms"]
    print(sys.argv[2], "value", d["value"], "e2e", d["e2e"]["value"], "| acc_g1 end", t["msm_accumulate_g1"]["end"], "reduce end", t["msm_reduce"]["end"], "span", t["_span"])
except Exception as e:
    print(sys.argv[2], "FAILED", e)
PY
}
for n in 8 4; do for f in 3 -1; do for o in "reduce_l=16" "reduce_l=8" "reduce_l=4" "reduce_l=8 tree_threads=32" "reduce_l=4 tree_threads=32" "reduce_l=8 reduce_l_g2=4"; do
  timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --emulate-shards $n --emulate-poly-mask 1 --opt fuse_g1=$f $o > gpurun_out/sweep.json 2> gpurun_out/sweep.log
  show gpurun_out/sweep.json "emu$n fuse=$f $o"
done; done; done
