#!/bin/bash
# Round-2 end measurement on one B200: both bench arms, ncu launch list, ncu --set full of the accumulation and
# reduction kernels (2^20), the 2^24 configuration, the single-GPU MSM sweep with the reference's CPU multiexp
mkdir -p gpurun_out
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_reference.json 2> gpurun_out/r02_bench_reference.log
tail -c 500 gpurun_out/r02_bench_reference.json
timeout 400 python bench.py > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.log
tail -c 2500 gpurun_out/r02_bench_n1.json; tail -3 gpurun_out/r02_bench_n1.log
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
timeout 400 python tools/msm_bench.py --log-n 16 18 20 22 24 26 --iters 3 --cpu-max-log-n 24 > gpurun_out/r02_msm_sweep_g1_n1_cpu.jsonl 2> gpurun_out/msm_n1.log
timeout 200 python tools/msm_bench.py --log-n 20 24 --iters 3 --scalars fr > gpurun_out/r02_msm_sweep_g1_n1_fr.jsonl 2>> gpurun_out/msm_n1.log
timeout 200 python tools/msm_bench.py --log-n 20 24 --iters 3 --g2 > gpurun_out/r02_msm_sweep_g2_n1.jsonl 2>> gpurun_out/msm_n1.log
cut -c1-230 gpurun_out/r02_msm_sweep_g1_n1_cpu.jsonl gpurun_out/r02_msm_sweep_g1_n1_fr.jsonl gpurun_out/r02_msm_sweep_g2_n1.jsonl
