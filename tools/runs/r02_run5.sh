#!/bin/bash
# round 2, GPU call 5: sharded path after the head start / slice combine / fused launch changes (emulated rank 0 of N)
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_benchsize.py -m gpu -x -q -k "shard or exchange or spread or stage" 2>&1 | tail -4 )
show() { python - "$1" <<'PY'
import json,sys
try:
    d=json.loads([l for l in open(sys.argv[1]) if l.startswith("{")][-1])
    t=d["timeline_ms"]
    print(sys.argv[1].split("/")[-1], "value", d["value"], "e2e", d["e2e"]["value"],
          "| sort", t.get("msm_sort"), "acc_g2", t.get("msm_accumulate_g2"), "acc_g1", t.get("msm_accumulate_g1"), "merge", t.get("msm_merge"), "reduce", t.get("msm_reduce"), "ntt", t.get("ntt_h"), "span", t["_span"])
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
}
for n in 8 4 2; do for f in -1 3 0; do
  timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --emulate-shards $n --emulate-poly-mask 1 --opt fuse_g1=$f > gpurun_out/r02b_fuse${f}_emu$n.json 2> gpurun_out/r02b_fuse${f}_emu$n.log
  show gpurun_out/r02b_fuse${f}_emu$n.json
done; done
B200_TIMELINE=1 timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --emulate-shards 8 --emulate-poly-mask 1 2>&1 | grep timeline | tail -24
