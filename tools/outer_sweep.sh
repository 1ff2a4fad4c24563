#!/bin/bash
# Outer-block width of the factorisation / solves (SPDE_FACTOR_OUTER, SPDE_SOLVE_OUTER; default 8 blocks = 512 columns) on C3.
mkdir -p gpurun_out
for o in 4 16 8; do
  SPDE_FACTOR_OUTER=$o timeout 70 python bench.py --workload c3 --steps 3 --warmup 3 --no-cpu > gpurun_out/r1_outer_$o.json 2> gpurun_out/r1_outer_$o.err
  python -c "
import json
d=json.loads(open('gpurun_out/r1_outer_$o.json').read().strip().splitlines()[-1]); print('outer', $o, d['value'], d['ms_per_step'], d['cholesky_gflops'])
"
done
