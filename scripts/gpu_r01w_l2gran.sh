mkdir -p gpurun_out
scripts/micro/gather > gpurun_out/r01w_gather_micro_l2gran.txt 2>&1; grep -E "default|set to|268 MB, 8|1024 MB, 8" gpurun_out/r01w_gather_micro_l2gran.txt
PL_L2=0,128,64,32 PL_VARIANTS=0,4 PL_LOG2=25 PL_MODES="ap[dp_sp_hp]" PL_SPLITS=256 timeout 400 python scripts/tune_ap.py 2>&1 | tail -14
