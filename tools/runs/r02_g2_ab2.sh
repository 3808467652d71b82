#!/bin/bash
# Second G2 experiment: the 168-register G2 accumulation (3 CTAs = 12 warps per SM) with the next point prefetched
# into L2 instead of registers (default build), with the sequential Y3 (g2seq build) and with an L1 prefetch (g2pfl1),
# each beside lockstep_g1; then the best candidates at 2^22 and 2^24.
mkdir -p gpurun_out
M=smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio,smsp__average_warps_issue_stalled_wait_per_issue_active.ratio,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio,sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed,smsp__issue_active.avg.pct_of_peak_sustained_active,gpu__time_duration.sum,sm__warps_active.avg.pct_of_peak_sustained_active
run() { # tag, lib, log_n, iters, configs...
    tag=$1; lib=$2; ln=$3; it=$4; shift 4
    B200_LIB=$lib timeout 400 python tools/prove_bench.py --log-n $ln --iters $it --configs "$@" > gpurun_out/r02_g2_ab2_$tag.jsonl 2> gpurun_out/r02_g2_ab2_$tag.log
    cut -c1-330 gpurun_out/r02_g2_ab2_$tag.jsonl; tail -2 gpurun_out/r02_g2_ab2_$tag.log | cut -c1-200
}
D=$PWD/rapidsnark_old_b200/libb200snark.so
run main $D 20 5 "" lockstep_g1=1 g2_minb=3 g2_minb=3,lockstep_g1=1 g2_minb=3,lockstep_g1=1,lockstep_g2=1 reduce_l_tail=8 reduce_l_tail=4 reduce_l_tail=8,lockstep_g1=1 ""
run seq $PWD/build/ab/libb200snark_g2seq.so 20 5 g2_minb=3 g2_minb=3,lockstep_g1=1 g2_minb=2
run pfl1 $PWD/build/ab/libb200snark_g2pfl1.so 20 5 g2_minb=3 g2_minb=3,lockstep_g1=1
timeout 300 ncu --metrics $M --clock-control none -k 'regex:k_msm_accumulate<' --csv --log-file gpurun_out/r02_g2_ab2_ncu.csv \
    python tools/prove_bench.py --log-n 20 --iters 1 --configs g2_minb=3 > gpurun_out/r02_g2_ab2_ncu.log 2>&1
B200_LIB=$PWD/build/ab/libb200snark_g2seq.so timeout 300 ncu --metrics $M --clock-control none -k 'regex:k_msm_accumulate<' --csv --log-file gpurun_out/r02_g2_ab2_ncu_seq.csv \
    python tools/prove_bench.py --log-n 20 --iters 1 --configs g2_minb=3 > gpurun_out/r02_g2_ab2_ncu_seq.log 2>&1
run m22 $D 22 3 "" lockstep_g1=1 g2_minb=3,lockstep_g1=1
run m24 $D 24 2 "" lockstep_g1=1 g2_minb=3,lockstep_g1=1
