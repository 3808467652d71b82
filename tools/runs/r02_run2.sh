#!/bin/bash
# round 2, GPU call 2: DFMA microbenchmark; fast squaring A/B
mkdir -p gpurun_out
timeout 300 build/dfma_ubench > gpurun_out/r02_dfma_ubench.txt 2>&1; cat gpurun_out/r02_dfma_ubench.txt
for v in default nosqr; do
  lib=""; [ $v = nosqr ] && lib="B200_LIB=$PWD/build/ab/libb200snark_nosqr.so"
  env $lib timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02_ab_$v.json 2> gpurun_out/r02_ab_$v.log
  python - gpurun_out/r02_ab_$v.json <<'PY'
import json,sys
d=json.loads([l for l in open(sys.argv[1]) if l.startswith("{")][-1])
print(sys.argv[1], d["value"], d["e2e"]["value"], d["circom_like_witness"]["value"], d["phase_stream_ms"]["msm_accumulate_g1"], d["phase_stream_ms"]["msm_accumulate_g2"])
PY
done
