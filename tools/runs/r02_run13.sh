#!/bin/bash
# round 2, GPU call 13: full GPU suite on the new kernels (CTA-aggregated task plan, two-lane bucket reduction,
# lockstep G1 launches), then A/B of the reduce segment lengths under the two-lane reduction, and one emulated shard
mkdir -p gpurun_out
( timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 ) | tee gpurun_out/r02_run13_pytest.log
timeout 400 python tools/prove_bench.py --log-n 20 --iters 5 --configs "" lockstep_g1=0 reduce_l_tail=32 reduce_l_tail=8 reduce_l=128 reduce_l=32 reduce_l_g2=128 "" \
    > gpurun_out/r02_run13_ab.jsonl 2> gpurun_out/r02_run13_ab.log
cut -c1-400 gpurun_out/r02_run13_ab.jsonl; tail -2 gpurun_out/r02_run13_ab.log
for r in 0 7; do
  timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --emulate-shards 8 --emulate-rank $r --emulate-poly-mask -2 > gpurun_out/r02_run13_emu8_r$r.json 2> gpurun_out/r02_run13_emu8_r$r.log
  python - gpurun_out/r02_run13_emu8_r$r.json <<'PY'
import json,sys
d=json.loads([l for l in open(sys.argv[1]) if l.startswith("{")][-1])
print(sys.argv[1], "value", d["value"], "span", d.get("timeline_ms", {}).get("_span"), {k: v.get("busy") for k, v in d.get("timeline_ms", {}).items() if isinstance(v, dict)})
PY
done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_run13_launches.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r02_run13_ncu.log 2>&1
tail -1 gpurun_out/r02_run13_ncu.log | cut -c1-300
