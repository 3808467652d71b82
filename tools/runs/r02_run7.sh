#!/bin/bash
# round 2, GPU call 7: uneven shard plan + L = 8 on emulated ranks of 8 / 4 / 2 (owner rank 0 and a non-owner rank)
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "shard or exchange or spread or stage or fused" 2>&1 | tail -3 )
show() { python - "$1" "$2" <<'PY'
import json,sys
try:
    d=json.loads([l for l in open(sys.argv[1]) if l.startswith("{")][-1])
    t=d["timeline_ms"]
    print(sys.argv[2], "value", d["value"], "e2e", d["e2e"]["value"], "| g2 end", t["msm_accumulate_g2"]["end"], "ntt end", t.get("ntt_h",{}).get("end"), "acc_g1", t["msm_accumulate_g1"]["start"], t["msm_accumulate_g1"]["end"], "reduce end", t["msm_reduce"]["end"], "span", t["_span"])
except Exception as e:
    print(sys.argv[2], "FAILED", e)
PY
}
for cfg in "8 0" "8 7" "4 0" "4 3" "2 0" "2 1"; do set -- $cfg; for ev in "" "--even-shards"; do
  timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --emulate-shards $1 --emulate-rank $2 --emulate-poly-mask -2 $ev > gpurun_out/sweep.json 2> gpurun_out/sweep.log
  show gpurun_out/sweep.json "emu$1 rank$2 $ev"
done; done
