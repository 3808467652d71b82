#!/bin/bash
# One 8-GPU box call: proof bench at N = 8 / 4 / 2 and the sharded G1 MSM sweep (results under gpurun_out/).
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
run() { timeout "$1" "${@:2}"; }
run 100 $TR --nproc-per-node 8 --master-port 29601 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r01_bench_n8.json 2> gpurun_out/bench_n8.log
run 100 $TR --nproc-per-node 8 --master-port 29602 bench.py --gpus 8 --steps 10 --warmup 3 --replicate-h > gpurun_out/r01_bench_n8_replicated_h.json 2> gpurun_out/bench_n8r.log
( CUDA_VISIBLE_DEVICES=0,1,2,3 run 100 $TR --nproc-per-node 4 --master-port 29603 bench.py --gpus 4 --steps 10 --warmup 3 > gpurun_out/r01_bench_n4.json 2> gpurun_out/bench_n4.log ) &
( CUDA_VISIBLE_DEVICES=4,5,6,7 run 100 $TR --nproc-per-node 4 --master-port 29604 tools/msm_bench.py --log-n 20 22 24 26 --iters 3 > gpurun_out/r01_msm_sweep_g1_n4.jsonl 2> gpurun_out/msm_n4.log ) &
wait
( CUDA_VISIBLE_DEVICES=0,1 run 100 $TR --nproc-per-node 2 --master-port 29605 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r01_bench_n2.json 2> gpurun_out/bench_n2.log ) &
( CUDA_VISIBLE_DEVICES=2,3 run 100 $TR --nproc-per-node 2 --master-port 29606 tools/msm_bench.py --log-n 20 22 24 26 --iters 3 > gpurun_out/r01_msm_sweep_g1_n2.jsonl 2> gpurun_out/msm_n2.log ) &
( CUDA_VISIBLE_DEVICES=4 run 100 python tools/msm_bench.py --log-n 16 18 20 22 24 --iters 3 --cpu-max-log-n 22 > gpurun_out/r01_msm_sweep_g1_n1_cpu.jsonl 2> gpurun_out/msm_n1.log ) &
wait
run 120 $TR --nproc-per-node 8 --master-port 29607 tools/msm_bench.py --log-n 20 22 24 26 28 --iters 3 > gpurun_out/r01_msm_sweep_g1_n8.jsonl 2> gpurun_out/msm_n8.log
if [ "$1" = "with26" ]; then
  # BASELINE config 4: 2^26 constraints sharded over 8 GPUs; every rank generates only its own table slices
  run 900 $TR --nproc-per-node 8 --master-port 29608 bench.py --gpus 8 --log-n 26 --steps 3 --warmup 3 > gpurun_out/bench_2_26_n8.json 2> gpurun_out/bench_2_26_n8.log
  tail -c 1200 gpurun_out/bench_2_26_n8.json; tail -3 gpurun_out/bench_2_26_n8.log
fi
for f in gpurun_out/r01_bench_n8.json gpurun_out/r01_bench_n8_replicated_h.json gpurun_out/r01_bench_n4.json gpurun_out/r01_bench_n2.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads([l for l in open(sys.argv[1]).read().splitlines() if l.startswith("{")][-1])
    print(sys.argv[1], d["n_gpus"], d["value"], d["e2e"]["value"], d["phases_ms"])
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
done
cut -c1-160 gpurun_out/r01_msm_sweep_g1_n*.jsonl
tail -2 gpurun_out/*.log | grep -iE "error|Traceback" | head
