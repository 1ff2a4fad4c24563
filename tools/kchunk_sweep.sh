for kc in 0 1024 512 256; do
  SPDE_SELINV_KCHUNK=$kc python bench.py --workload c3 --steps 3 --warmup 3 --no-cpu > gpurun_out/r1_kc_$kc.json 2> gpurun_out/r1_kc_$kc.err
  python -c "
import json
d=json.load(open('gpurun_out/r1_kc_$kc.json')); r=d['roofline']; print('kchunk', $kc, d['value'], d['ms_per_step'], r['by_kind_ms']['gemm'], r['launches_per_step'])
"
done
