#!/bin/bash
# round 2, GPU call 12: tail reduce-segment length A/B at N = 1; proof server test
mkdir -p gpurun_out
( timeout 300 python -m pytest tests/test_golden.py -m gpu -x -q -k "server or fullprover" 2>&1 | tail -3 )
for o in "reduce_l_tail=16" "reduce_l_tail=8" "reduce_l_tail=4" "reduce_l_tail=32" "fuse_g1=4 reduce_l_tail=8" "fuse_g1=4 reduce_l_tail=16"; do
  timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --opt $o > gpurun_out/sweep.json 2> gpurun_out/sweep.log
  python - gpurun_out/sweep.json "$o" <<'PY'
import json,sys
d=json.loads([l for l in open(sys.argv[1]) if l.startswith("{")][-1])
t=d["timeline_ms"]
print(sys.argv[2], "value", d["value"], "e2e", d["e2e"]["value"], "acc_g1 end", t["msm_accumulate_g1"]["end"], "reduce end", t["msm_reduce"]["end"], "span", t["_span"])
PY
done
