#!/bin/bash
# Round-end measurement on one B200: official bench line, ncu launch list, ncu --set full of the accumulation
# kernels (2^20), then the 2^24 configuration (bench line + a metrics-limited ncu pass over accumulate / NTT kernels).
mkdir -p gpurun_out
timeout 300 python bench.py > gpurun_out/r01_bench_n1.json 2> gpurun_out/bench_n1.log
tail -c 600 gpurun_out/r01_bench_n1.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r01_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_msm_accumulate -c 5 -f -o gpurun_out/r01_accumulate \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out/*.ncu-rep
if [ "$1" = "with24" ]; then
  timeout 500 python bench.py --log-n 24 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r01_bench_2_24_new.json 2> gpurun_out/bench_2_24.log
  tail -c 1500 gpurun_out/r01_bench_2_24_new.json
  M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__throughput.avg.pct_of_peak_sustained_elapsed,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,lts__t_sector_hit_rate.pct,smsp__issue_active.avg.pct_of_peak_sustained_active
  timeout 600 ncu --metrics $M --clock-control none -k 'regex:k_msm_accumulate|k_ntt_pass' -c 23 --csv --log-file gpurun_out/r01_2_24_kernels.csv \
      python bench.py --log-n 24 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_2_24.log 2>&1
  tail -3 gpurun_out/ncu_2_24.log
fi
