#!/bin/bash
# A/B: which part of the lean loop helps / hurts (dp / sp / hp, SELL-32, 7-point 256^3)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
L=ultimate-spmv_b200/lib/ab
{
USPMV_B200_LIB=$PWD/$L/libold.so python scripts/ab_stream.py old 0,1,2
USPMV_B200_LIB=$PWD/$L/libB_oldloop_tbody.so python scripts/ab_stream.py oldloop_templated_body 0,1,2
USPMV_B200_LIB=$PWD/$L/libA_newloop_rtbody.so python scripts/ab_stream.py newloop_runtime_body 0,13,14,1,2
USPMV_B200_LIB=$PWD/$L/libnew.so python scripts/ab_stream.py new 0,13,14,1,2,8
} 2>&1 | tee gpurun_out/r02d_ab_stream.log
