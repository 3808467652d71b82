#!/bin/bash
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 240 $TR --nproc-per-node 2 --master-port 29605 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r02_n2check.json 2> gpurun_out/r02_n2check.log
python - <<'PY'
import json
try:
    d=json.loads([l for l in open("gpurun_out/r02_n2check.json").read().splitlines() if l.startswith("{")][-1])
    print("N2", d["value"], d["e2e"]["value"], d["circom_like_witness"]["value"], d["timeline_ms"]["_span"], d.get("host_path_ms"))
except Exception as e:
    print("FAILED", e)
PY
tail -3 gpurun_out/r02_n2check.log
