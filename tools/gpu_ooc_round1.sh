#!/bin/bash
# GPU session for the streamed evaluator: full GPU suite, C3 through the streamed path (1 GB threshold), then the
# headline mesh 256x256x100 (BASELINE configs[3]).  Outputs under gpurun_out/.
mkdir -p gpurun_out
set -o pipefail
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/gpu_tests.log; rc=${PIPESTATUS[0]}
echo "pytest rc=$rc"
[ $rc -ne 0 ] && exit 1
SPDE_OOC_TOP_BYTES=1e9 timeout 600 python bench.py --workload c3 --streamed --steps 2 --warmup 1 --check-forward \
    > gpurun_out/c3_streamed.json 2> gpurun_out/c3_streamed.err; rc=$?
echo "c3 streamed rc=$rc"; tail -3 gpurun_out/c3_streamed.err; head -c 600 gpurun_out/c3_streamed.json; echo
[ $rc -ne 0 ] && exit 1
timeout 1500 python bench.py --workload c4 --steps 1 --warmup 0 > gpurun_out/c4.json 2> gpurun_out/c4.err; rc=$?
echo "c4 rc=$rc"; tail -5 gpurun_out/c4.err; head -c 1500 gpurun_out/c4.json; echo
nvidia-smi --query-gpu=memory.used,memory.total --format=csv
free -g | head -2
