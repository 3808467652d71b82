#!/bin/bash
# Round-2 end measurement on one B200 (after the CTA-aggregated task plan, the two-lane bucket reduction, lockstep G1
# launches, segment lengths 128 / 32): GPU suite, smoke, both bench arms, ncu launch list, ncu --set full of the
# accumulation and reduction kernels (2^20), the 2^24 configuration
mkdir -p gpurun_out
( timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 ) | tee gpurun_out/r02_final2_pytest.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee gpurun_out/r02_final2_smoke.log
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_reference.json 2> gpurun_out/r02_bench_reference.log
tail -c 400 gpurun_out/r02_bench_reference.json
timeout 400 python bench.py > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.log
tail -c 3000 gpurun_out/r02_bench_n1.json; tail -3 gpurun_out/r02_bench_n1.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
# (the report itself is tens of MB: only its raw-metrics page travels back - gpurun_out/ is limited to 64 MiB)
timeout 500 ncu --set full --clock-control none -k 'regex:k_msm_accumulate|k_msm_reduce_segments' -c 8 -f -o /tmp/r02_accumulate \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
ls -la /tmp/r02_accumulate.ncu-rep; tail -2 gpurun_out/ncu_full.log | cut -c1-200
ncu -i /tmp/r02_accumulate.ncu-rep --page raw --csv > gpurun_out/r02_accumulate_raw.csv 2> /dev/null
ls -la gpurun_out/r02_accumulate_raw.csv
timeout 600 python bench.py --log-n 24 --steps 3 --warmup 3 > gpurun_out/r02_bench_2_24.json 2> gpurun_out/r02_bench_2_24.log
tail -c 1500 gpurun_out/r02_bench_2_24.json
timeout 200 python tools/msm_bench.py --log-n 20 24 --iters 3 > gpurun_out/r02_msm_n1_final.jsonl 2> gpurun_out/msm_n1.log
cut -c1-230 gpurun_out/r02_msm_n1_final.jsonl
