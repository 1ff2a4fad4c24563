# A/B of the two-lane look-ahead factor schedule (run on the GPU box)
for la in 1 0; do
  SPDE_LOOKAHEAD=$la python bench.py --workload c3 --steps 3 --warmup 3 --no-cpu > gpurun_out/r1_la_$la.json 2> gpurun_out/r1_la_$la.err
  tail -2 gpurun_out/r1_la_$la.err
  python -c "
import json
d=json.load(open('gpurun_out/r1_la_$la.json')); r=d['roofline']; print('lookahead', $la, d['value'], d['ms_per_step'], d['cholesky_gflops'], r['launches_per_step'], d['last_like'])
"
done
