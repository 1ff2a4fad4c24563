#!/bin/bash
# GPU session for the overlapped panel traffic of the streamed evaluator: full GPU suite, C3 through the streamed path
# with the overlap on and off (same value expected), then the headline mesh 256x256x100 (BASELINE configs[3]).
mkdir -p gpurun_out
set -o pipefail
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/gpu_tests_overlap.log; rc=${PIPESTATUS[0]}
echo "pytest rc=$rc"
[ $rc -ne 0 ] && exit 1
for ov in 1 0; do
  SPDE_OOC_OVERLAP=$ov SPDE_OOC_TOP_BYTES=2e8 timeout 600 python bench.py --workload c3 --streamed --steps 3 --warmup 2 --no-cpu \
      > gpurun_out/c3_streamed_ov$ov.json 2> gpurun_out/c3_streamed_ov$ov.err; rc=$?
  echo "c3 streamed overlap=$ov rc=$rc"; tail -2 gpurun_out/c3_streamed_ov$ov.err
  python - <<PY
import json
d = json.loads(open("gpurun_out/c3_streamed_ov$ov.json").read().strip().splitlines()[-1])
print("ms/step", d["ms_per_step"], "like", d.get("last_like"), "streamed", d.get("streamed"))
PY
  [ $rc -ne 0 ] && exit 1
done
timeout 1500 python bench.py --workload c4 --steps 1 --warmup 1 --no-cpu > gpurun_out/c4_overlap.json 2> gpurun_out/c4_overlap.err; rc=$?
echo "c4 rc=$rc"; tail -5 gpurun_out/c4_overlap.err; head -c 300 gpurun_out/c4_overlap.json; echo
