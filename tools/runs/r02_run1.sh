#!/bin/bash
# round 2, GPU call 1: parity suite (incl. the benchmark-size tests), bench N=1 both arms, launch list, 2^24 record
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv,noheader | head -2
( timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 ) > gpurun_out/r02_run1_pytest.log
tail -5 gpurun_out/r02_run1_pytest.log
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_reference.json 2> gpurun_out/r02_bench_reference.log
tail -c 600 gpurun_out/r02_bench_reference.json
timeout 400 python bench.py --steps 10 --warmup 3 > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.log
tail -c 3000 gpurun_out/r02_bench_n1.json; tail -5 gpurun_out/r02_bench_n1.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r02_launches.csv \
   python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r02_ncu_bench.log 2>&1
tail -2 gpurun_out/r02_ncu_bench.log | cut -c1-300
timeout 900 python bench.py --log-n 24 --steps 3 --warmup 3 > gpurun_out/r02_bench_2_24.json 2> gpurun_out/r02_bench_2_24.log
tail -c 2500 gpurun_out/r02_bench_2_24.json; tail -3 gpurun_out/r02_bench_2_24.log
