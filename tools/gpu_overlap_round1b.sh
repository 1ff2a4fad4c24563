#!/bin/bash
# Second GPU session for the overlapped panel traffic (schedule tables staged before the slices are fetched): streamed
# tests, C3 through the streamed path, then the headline mesh 256x256x100 (BASELINE configs[3]).
mkdir -p gpurun_out
set -o pipefail
timeout 600 python -m pytest tests/test_gpu_ooc.py -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/gpu_tests_overlap_b.log; rc=${PIPESTATUS[0]}
echo "pytest rc=$rc"
[ $rc -ne 0 ] && exit 1
SPDE_OOC_TOP_BYTES=2e8 timeout 600 python bench.py --workload c3 --streamed --steps 3 --warmup 2 --no-cpu \
    > gpurun_out/c3_streamed_ov1b.json 2> gpurun_out/c3_streamed_ov1b.err; rc=$?
echo "c3 streamed rc=$rc"; tail -1 gpurun_out/c3_streamed_ov1b.err
[ $rc -ne 0 ] && exit 1
timeout 1500 python bench.py --workload c4 --steps 1 --warmup 1 --no-cpu > gpurun_out/c4_overlap_b.json 2> gpurun_out/c4_overlap_b.err; rc=$?
echo "c4 rc=$rc"; tail -3 gpurun_out/c4_overlap_b.err; head -c 300 gpurun_out/c4_overlap_b.json; echo
