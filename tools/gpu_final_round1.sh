#!/bin/bash
# Last GPU session of round 1: full GPU suite on the final tree, launch profile of C3 with the corrected flop
# accounting of lower (trapezoid) GEMM tasks, compute-sanitizer memcheck of the smoke evaluation.
mkdir -p gpurun_out
set -o pipefail
timeout 150 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee gpurun_out/gpu_tests_final.log
echo "pytest rc=${PIPESTATUS[0]}"
timeout 60 python tools/launch_profile.py c3 > gpurun_out/launch_profile_c3.txt 2> gpurun_out/launch_profile_c3.err; echo "profile rc=$?"
timeout 70 compute-sanitizer --tool memcheck --error-exitcode 3 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/sanitizer_smoke.txt 2>&1; echo "memcheck rc=$?"
tail -3 gpurun_out/sanitizer_smoke.txt
