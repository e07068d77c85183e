#!/bin/bash
# round 2, call i (1 GPU): ncu of the headline kernel and of the config-4 AP kernel (summaries only), launch list of the bench command,
# segment-length sweep of the AP kernel
cd "$(dirname "$0")/.."
mkdir -p gpurun_out /tmp/ncu
timeout 600 python scripts/ap_one.py 25 64,128,256,512,1024,0 > gpurun_out/r02i_ap_segment_sweep.log 2>&1; cat gpurun_out/r02i_ap_segment_sweep.log
timeout 900 ncu --set full --clock-control none -k regex:k_scs32_stream_ap --launch-skip 2 --launch-count 1 -o /tmp/ncu/ap -f python scripts/ap_one.py 25 > /tmp/ncu/ap.log 2>&1; echo "ncu ap rc=$?"
python scripts/ncu_summary.py /tmp/ncu/ap.ncu-rep > gpurun_out/r02i_config4_ap_sigma512_ncu_summary.txt 2>&1
ncu -i /tmp/ncu/ap.ncu-rep --page details --csv > gpurun_out/r02i_config4_ap_sigma512_details.csv 2>/dev/null
timeout 600 ncu --set full --clock-control none -k regex:k_scs32_stream --launch-skip 3 --launch-count 1 -o /tmp/ncu/dp -f python scripts/ab_one.py dp > /tmp/ncu/dp.log 2>&1; echo "ncu dp rc=$?"
python scripts/ncu_summary.py /tmp/ncu/dp.ncu-rep > gpurun_out/r02i_spmv_dp_ncu_summary.txt 2>&1
ncu -i /tmp/ncu/dp.ncu-rep --page details --csv > gpurun_out/r02i_spmv_dp_details.csv 2>/dev/null
python bench.py --steps 2 --warmup 1 --no-other-configs --no-cpu-baseline --no-e2e --steady-steps 0 > /tmp/ncu/b.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02i_bench_kernel_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-other-configs --no-cpu-baseline --no-e2e --steady-steps 0 > /tmp/ncu/ncu_b.log 2>&1; echo "launch list rc=$?"
ls -la gpurun_out/r02i_*; cat gpurun_out/r02i_config4_ap_sigma512_ncu_summary.txt | cut -c1-140
