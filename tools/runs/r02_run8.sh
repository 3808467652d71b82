#!/bin/bash
# round 2, GPU call 8: full GPU suite after the ingestion rewrite; CLI wall clock at 2^20
mkdir -p gpurun_out
free -g | head -2; nproc
( timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 )
timeout 600 python tools/cli_bench.py --log-n 20 > gpurun_out/r02_cli_bench.json 2> gpurun_out/r02_cli_bench.log; cat gpurun_out/r02_cli_bench.json; grep -E "gpu0:|context" gpurun_out/r02_cli_bench.log | tail -3
B200_NO_STAGING=1 timeout 600 python tools/cli_bench.py --log-n 20 --skip-reference 2>&1 | tail -2
