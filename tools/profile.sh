#!/bin/bash
# Profiling recipe (run under gpurun, ONE GPU): launch list of one UNet step + ncu --set full captures of
# the two dominant kernels. Outputs land in gpurun_out/ (copy summaries into profiles/ afterwards).
set -u
TAG=${1:-r01}
mkdir -p gpurun_out
BENCH="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-kernel-roofline"
# 1) every launch of ~1.2 UNet steps with its device time (skip init + encode + warm-up launches)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 1700 -c 560 --csv \
    --log-file gpurun_out/${TAG}_launches.csv $BENCH > gpurun_out/${TAG}_launches.log 2>&1
# 2) top kernels, full sections (a 96x96 conv tile kernel and the 9216-token flash attention)
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 330 -c 2 \
    -o gpurun_out/${TAG}_gemm -f $BENCH > gpurun_out/${TAG}_gemm.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:flash_attn64 -s 20 -c 1 \
    -o gpurun_out/${TAG}_attn -f $BENCH > gpurun_out/${TAG}_attn.log 2>&1
ls -la gpurun_out/
