# SpMMV sweep + ncu of the power-law kernels (one GPU; run under gpurun)
mkdir -p gpurun_out
python scripts/tune_mmv2.py > gpurun_out/r01j_tune_mmv2.log 2>&1
tail -30 gpurun_out/r01j_tune_mmv2.log
python scripts/one_powerlaw.py 22 16384 > gpurun_out/r01j_powerlaw_plain.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_scs32_stream_split|k_scs32_stream_ap' -s 2 -c 1 -f -o gpurun_out/r01j_powerlaw_split python scripts/one_powerlaw.py 22 16384 > gpurun_out/r01j_ncu1.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_scs32_stream_ap' -s 2 -c 1 -f -o gpurun_out/r01j_powerlaw_ap python scripts/one_powerlaw.py 22 16384 > gpurun_out/r01j_ncu2.log 2>&1
tail -3 gpurun_out/r01j_powerlaw_plain.log gpurun_out/r01j_ncu1.log gpurun_out/r01j_ncu2.log
