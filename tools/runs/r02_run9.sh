#!/bin/bash
# diagnose context creation time
nvidia-smi --query-gpu=persistence_mode,name --format=csv,noheader
cat > /tmp/ctx.cu <<'CU'
#include <cuda_runtime.h>
#include <chrono>
#include <cstdio>
int main() { auto t0 = std::chrono::steady_clock::now(); cudaFree(0); auto t1 = std::chrono::steady_clock::now();
  printf("bare cudaFree(0): %.1f ms\n", std::chrono::duration<double, std::milli>(t1 - t0).count()); return 0; }
CU
nvcc -o /tmp/ctx /tmp/ctx.cu && /tmp/ctx && /tmp/ctx
for i in 1 2; do python - <<'PY'
import time, sys
sys.path.insert(0, ".")
t=time.perf_counter()
import rapidsnark_old_b200 as b200
c=b200.Context(0)
print("python: import + Context(0): %.1f ms" % ((time.perf_counter()-t)*1e3))
PY
done
python - <<'PY'
import sys, os, subprocess, time, tempfile
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import rapidsnark_old_b200 as b200
from rapidsnark_old_b200 import synth
import bench
ctx = b200.Context(0)
s = bench.build_inputs(16, 2, *bench.gpu_point_makers(ctx))
d = tempfile.mkdtemp()
open(d+"/c.zkey","wb").write(synth.zkey_bytes(s)); open(d+"/w.wtns","wb").write(synth.wtns_bytes_file(s))
for env in ({}, {"CUDA_MODULE_LOADING": "LAZY"}, {"CUDA_MODULE_LOADING": "EAGER"}, {"CUDA_DEVICE_MAX_CONNECTIONS": "8"}):
    for rep in range(2):
        t=time.perf_counter()
        r=subprocess.run(["build/prover", d+"/c.zkey", d+"/w.wtns", d+"/p.json", d+"/pub.json"], env=dict(os.environ, B200_TIMING="1", **env), capture_output=True, text=True)
        print(env, "wall %.3f s" % (time.perf_counter()-t), r.stderr.strip().splitlines()[:2])
PY
