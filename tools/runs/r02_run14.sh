#!/bin/bash
# round 2, GPU call 14: segment lengths and plane-sum chunking under the two-lane bucket reduction - one GPU whole
# proof, then an emulated rank of 8 (rank 7 is the slowest: it owns no chain and the largest point range)
mkdir -p gpurun_out
timeout 400 python tools/prove_bench.py --log-n 20 --iters 5 --configs "" plane_items=8 reduce_l=128 reduce_l=128,reduce_l_tail=32 reduce_l=256,reduce_l_tail=32 \
    reduce_l=128,reduce_l_tail=64 reduce_l=128,reduce_l_g2=256,reduce_l_tail=32 reduce_l=128,reduce_l_tail=32,plane_items=4 "" \
    > gpurun_out/r02_run14_ab.jsonl 2> gpurun_out/r02_run14_ab.log
cut -c1-330 gpurun_out/r02_run14_ab.jsonl; tail -2 gpurun_out/r02_run14_ab.log
emu() { # tag rank opts...
  tag=$1; r=$2; shift 2
  timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --emulate-shards 8 --emulate-rank $r --emulate-poly-mask -2 --opt "$@" > gpurun_out/r02_run14_emu8_$tag.json 2> gpurun_out/r02_run14_emu8_$tag.log
  python - gpurun_out/r02_run14_emu8_$tag.json <<'PY'
import json,sys
try:
    d=json.loads([l for l in open(sys.argv[1]) if l.startswith("{")][-1])
    t=d.get("timeline_ms", {})
    print(sys.argv[1], "value", d["value"], "span", t.get("_span"), "acc_g1 end", t.get("msm_accumulate_g1", {}).get("end"), "reduce end", t.get("msm_reduce", {}).get("end"))
except Exception as e:
    print(sys.argv[1], "failed", e)
PY
}
emu r7_default 7 timeline=0
emu r7_l4 7 reduce_l=4
emu r7_l16 7 reduce_l=16
emu r7_items8 7 plane_items=8
emu r7_g2l4 7 reduce_l_g2=4
emu r0_default 0 timeline=0
emu r0_l4 0 reduce_l=4
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_run14_emu8_r7_launches.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --emulate-shards 8 --emulate-rank 7 --emulate-poly-mask -2 > gpurun_out/r02_run14_ncu.log 2>&1
tail -1 gpurun_out/r02_run14_ncu.log | cut -c1-200
