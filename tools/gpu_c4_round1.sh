#!/bin/bash
# Headline mesh 256x256x100 (BASELINE configs[3]): one warm-up evaluation (one-time allocations: 157 GB pool, 126 GB pinned
# host), one timed evaluation, one profiled evaluation (per-launch events).
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_ooc.py -x -q -k loglike 2>&1 | tail -3
timeout 1600 python bench.py --workload c4 --steps 1 --warmup 1 --profile-step > gpurun_out/c4_steady.json 2> gpurun_out/c4_steady.err; rc=$?
echo "c4 rc=$rc"; tail -5 gpurun_out/c4_steady.err; head -c 400 gpurun_out/c4_steady.json; echo
