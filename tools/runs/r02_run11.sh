#!/bin/bash
# round 2, GPU call 11: full GPU suite; N = 1 A/B of the H-pipeline start (after / beside the witness sort)
mkdir -p gpurun_out
( timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 ) | tee gpurun_out/r02_run11_pytest.log
for h in 0 1; do
  timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --opt h_early=$h > gpurun_out/r02_hearly$h.json 2> gpurun_out/r02_hearly$h.log
  python - gpurun_out/r02_hearly$h.json <<'PY'
import json,sys
d=json.loads([l for l in open(sys.argv[1]) if l.startswith("{")][-1])
t=d["timeline_ms"]
print(sys.argv[1], "value", d["value"], "e2e", d["e2e"]["value"], d["e2e"]["pageable_host_witness_ms"], "circom", d["circom_like_witness"]["value"], "sort", t["msm_sort"]["start"], "g2", t["msm_accumulate_g2"]["start"], t["msm_accumulate_g2"]["end"], "ntt end", t["ntt_h"]["end"], "span", t["_span"])
PY
done
