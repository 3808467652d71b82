#!/bin/bash
# round 2, GPU call 3: new tests (fused prove, views), bench with the fused call, throughput mode, sqr ubench A/B
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "fused or views or two_stage or in_process" 2>&1 | tail -5 )
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02_fused_n1.json 2> gpurun_out/r02_fused_n1.log
python - <<'PY'
import json
d=json.loads([l for l in open("gpurun_out/r02_fused_n1.json") if l.startswith("{")][-1])
print("bench", d["value"], d["e2e"], d["circom_like_witness"]["value"], d["timeline_ms"]["_span"])
PY
tail -3 gpurun_out/r02_fused_n1.log
timeout 300 python tools/throughput_bench.py --contexts 1 2 3 --seconds 3 > gpurun_out/r02_throughput_n1.jsonl 2> gpurun_out/r02_throughput_n1.log
cat gpurun_out/r02_throughput_n1.jsonl | cut -c1-260; tail -2 gpurun_out/r02_throughput_n1.log
timeout 120 build/dfma_ubench_nosqr 2>&1 | grep -E "mask= 0|products  fp_mul \(IMAD" 
