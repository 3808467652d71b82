#!/bin/bash
# G2 accumulation experiments on one B200 (2^20 proof, per-phase device times, five result points compared):
# lockstep warps (options lockstep_g1 / lockstep_g2) and the out-of-line Fq2 products (B200_G2_HOT_CALLS=3 build)
mkdir -p gpurun_out
M=smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio,smsp__average_warps_issue_stalled_wait_per_issue_active.ratio,sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed,smsp__issue_active.avg.pct_of_peak_sustained_active,gpu__time_duration.sum,sm__warps_active.avg.pct_of_peak_sustained_active
timeout 300 python tools/prove_bench.py --log-n 20 --iters 5 --configs "" lockstep_g2=1 lockstep_g2=2 lockstep_g1=1 lockstep_g1=1,lockstep_g2=1 lockstep_g1=1,lockstep_g2=2 "" \
    > gpurun_out/r02_g2_ab_default.jsonl 2> gpurun_out/r02_g2_ab_default.log
cut -c1-420 gpurun_out/r02_g2_ab_default.jsonl; tail -3 gpurun_out/r02_g2_ab_default.log
B200_LIB=$PWD/build/ab/libb200snark_g2calls3.so timeout 300 python tools/prove_bench.py --log-n 20 --iters 5 \
    --configs g2_minb=2 g2_minb=3 g2_minb=2,lockstep_g2=1 g2_minb=2,lockstep_g2=2 g2_minb=3,lockstep_g2=1 \
    > gpurun_out/r02_g2_ab_calls3.jsonl 2> gpurun_out/r02_g2_ab_calls3.log
cut -c1-420 gpurun_out/r02_g2_ab_calls3.jsonl; tail -3 gpurun_out/r02_g2_ab_calls3.log
timeout 300 ncu --metrics $M --clock-control none -k 'regex:k_msm_accumulate' --csv --log-file gpurun_out/r02_g2_ab_ncu.csv \
    python tools/prove_bench.py --log-n 20 --iters 1 --configs "" lockstep_g2=1 lockstep_g2=2 lockstep_g1=1 > gpurun_out/r02_g2_ab_ncu.log 2>&1
tail -2 gpurun_out/r02_g2_ab_ncu.log | cut -c1-300
B200_LIB=$PWD/build/ab/libb200snark_g2calls3.so timeout 300 ncu --metrics $M --clock-control none -k 'regex:k_msm_accumulate' --csv --log-file gpurun_out/r02_g2_ab_ncu_calls3.csv \
    python tools/prove_bench.py --log-n 20 --iters 1 --configs g2_minb=2 g2_minb=2,lockstep_g2=1 > gpurun_out/r02_g2_ab_ncu_calls3.log 2>&1
timeout 300 python tools/prove_bench.py --log-n 22 --iters 3 --configs "" lockstep_g2=1 lockstep_g2=2 \
    > gpurun_out/r02_g2_ab_2_22.jsonl 2> gpurun_out/r02_g2_ab_2_22.log
cut -c1-420 gpurun_out/r02_g2_ab_2_22.jsonl; tail -3 gpurun_out/r02_g2_ab_2_22.log
