#!/bin/bash
# round 2, the 8-GPU call: proof bench at N = 8 / 4 / 2, in-process 8-GPU CLI vs the reference CLI (byte identity),
# BASELINE config 4 (2^26 constraints over 8 GPUs), sharded G1 MSM sweep, 8 replicas in throughput mode
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
run() { timeout "$1" "${@:2}"; }
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8; free -g | head -2; nproc
show() { python - "$1" <<'PY'
import json,sys
try:
    d=json.loads([l for l in open(sys.argv[1]).read().splitlines() if l.startswith("{")][-1])
    t=d.get("timeline_ms",{})
    print(sys.argv[1].split("/")[-1], "N", d.get("n_gpus"), "value", d.get("value"), "e2e", (d.get("e2e") or {}).get("value"), "circom", (d.get("circom_like_witness") or {}).get("value"), "span", t.get("_span"), d.get("skipped"))
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
}
run 200 $TR --nproc-per-node 8 --master-port 29601 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r02_bench_n8.json 2> gpurun_out/r02_bench_n8.log; show gpurun_out/r02_bench_n8.json
run 200 $TR --nproc-per-node 8 --master-port 29611 bench.py --gpus 8 --steps 10 --warmup 3 --even-shards > gpurun_out/r02_bench_n8_even.json 2> gpurun_out/r02_bench_n8_even.log; show gpurun_out/r02_bench_n8_even.json
( CUDA_VISIBLE_DEVICES=0,1,2,3 run 200 $TR --nproc-per-node 4 --master-port 29603 bench.py --gpus 4 --steps 10 --warmup 3 > gpurun_out/r02_bench_n4.json 2> gpurun_out/r02_bench_n4.log ) &
( CUDA_VISIBLE_DEVICES=4,5 run 200 $TR --nproc-per-node 2 --master-port 29605 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r02_bench_n2.json 2> gpurun_out/r02_bench_n2.log ) &
( CUDA_VISIBLE_DEVICES=6,7 run 200 $TR --nproc-per-node 2 --master-port 29606 tools/msm_bench.py --log-n 20 22 24 26 --iters 3 > gpurun_out/r02_msm_sweep_g1_n2.jsonl 2> gpurun_out/msm_n2.log ) &
wait
show gpurun_out/r02_bench_n4.json; show gpurun_out/r02_bench_n2.json
run 300 python tools/cli_bench.py --log-n 20 --gpus 8 > gpurun_out/r02_cli_bench_8gpu.json 2> gpurun_out/r02_cli_bench_8gpu.log; cat gpurun_out/r02_cli_bench_8gpu.json; tail -2 gpurun_out/r02_cli_bench_8gpu.log
run 600 $TR --nproc-per-node 8 --master-port 29608 bench.py --gpus 8 --log-n 26 --steps 3 --warmup 3 > gpurun_out/r02_bench_2_26_n8.json 2> gpurun_out/r02_bench_2_26_n8.log; show gpurun_out/r02_bench_2_26_n8.json; grep -E "host memory|Error|error" gpurun_out/r02_bench_2_26_n8.log | head -5
nvidia-smi --query-gpu=index,memory.used --format=csv,noheader | head -8
run 200 $TR --nproc-per-node 8 --master-port 29607 tools/msm_bench.py --log-n 20 22 24 26 28 --iters 3 > gpurun_out/r02_msm_sweep_g1_n8.jsonl 2> gpurun_out/msm_n8.log
run 100 $TR --nproc-per-node 8 --master-port 29609 tools/msm_bench.py --log-n 24 26 --iters 3 --scalars fr > gpurun_out/r02_msm_sweep_g1_n8_fr.jsonl 2> gpurun_out/msm_n8fr.log
( CUDA_VISIBLE_DEVICES=0,1,2,3 run 150 $TR --nproc-per-node 4 --master-port 29613 tools/msm_bench.py --log-n 20 22 24 26 --iters 3 > gpurun_out/r02_msm_sweep_g1_n4.jsonl 2> gpurun_out/msm_n4.log ) &
( CUDA_VISIBLE_DEVICES=4,5,6,7 run 150 python tools/throughput_bench.py --gpus 4 --contexts 2 --seconds 3 > gpurun_out/r02_throughput_4replicas.jsonl 2> gpurun_out/thr4.log ) &
wait
run 150 python tools/throughput_bench.py --gpus 8 --contexts 2 --seconds 3 > gpurun_out/r02_throughput_8replicas.jsonl 2> gpurun_out/thr8.log
cut -c1-200 gpurun_out/r02_msm_sweep_g1_n*.jsonl gpurun_out/r02_throughput_*replicas.jsonl
grep -liE "error|Traceback" gpurun_out/*.log | head
