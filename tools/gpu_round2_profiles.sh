#!/bin/bash
# Round-2 profiling session (one B200): ncu DRAM-traffic pass over every dense-tile launch of one C3 evaluation, full ncu captures of
# one big k_gemm_ws launch of the Takahashi schedule and one of the factorisation, the ncu launch list of the bench command, the
# per-kernel roofline table, launch profiles, and the C1 / C2 / Hutchinson bench lines.
mkdir -p gpurun_out
NCU="ncu --clock-control none"
timeout 900 $NCU --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --profile-from-start off -k regex:k_gemm \
    --csv --log-file gpurun_out/r2_gemm_traffic_c3.csv python tools/ncu_hbm_target.py c3 > gpurun_out/r2_gemm_traffic_c3.log 2>&1
python tools/ncu_gemm_traffic.py gpurun_out/r2_gemm_traffic_c3.csv > gpurun_out/r2_gemm_traffic_c3.json
timeout 300 $NCU --set full --import-source on --profile-from-start off -k regex:k_gemm_ws -s 10 -c 1 -o gpurun_out/r2_gemm_ws_selinv -f \
    python tools/ncu_target.py c3 selinv > gpurun_out/r2_ncu_ws_selinv.log 2>&1
timeout 300 $NCU --set full --import-source on --profile-from-start off -k regex:k_gemm_ws -s 40 -c 1 -o gpurun_out/r2_gemm_ws_factor -f \
    python tools/ncu_target.py c3 factor > gpurun_out/r2_ncu_ws_factor.log 2>&1
timeout 600 $NCU --metrics gpu__time_duration.sum -c 400 --csv --log-file gpurun_out/r2_ncu_launches_bench.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu --no-c4 --no-c5 > gpurun_out/r2_ncu_launches_bench.log 2>&1
python tools/kernel_roofline.py c3 > gpurun_out/r2_kernel_roofline_c3.txt 2> gpurun_out/r2_kernel_roofline_c3.err
python tools/launch_profile.py c3 > gpurun_out/r2_launch_profile_c3.txt 2> gpurun_out/r2_launch_profile_c3.err
python tools/launch_profile.py c2 > gpurun_out/r2_launch_profile_c2.txt 2> gpurun_out/r2_launch_profile_c2.err
python tools/timing_breakdown.py c2 > gpurun_out/r2_timing_breakdown_c2.txt 2>&1
python bench.py --workload c2 --steps 20 --warmup 5 > gpurun_out/r2_bench_c2.json 2> gpurun_out/r2_bench_c2.err
python bench.py --workload c1 --steps 50 --warmup 5 > gpurun_out/r2_bench_c1.json 2> gpurun_out/r2_bench_c1.err
python bench.py --gradient hutchinson --steps 3 --warmup 3 --no-cpu > gpurun_out/r2_bench_c3_hutchinson.json 2> gpurun_out/r2_bench_c3_hutchinson.err
tail -n 3 gpurun_out/*.err gpurun_out/r2_ncu_ws_*.log
ls -la gpurun_out | tail -n 30
