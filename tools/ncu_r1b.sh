#!/bin/bash
# ncu evidence for the second half of round 1 (run under gpurun, one GPU).  Every profiled command first runs to completion
# without ncu (rule: never profile a program that has not exited 0 on its own).
set -x
mkdir -p gpurun_out
# 1. launch list of the bench command at its full size (N=65536).  NOTE: ncu manages ~10 launches/s at this size (17 GB
#    buffers): the first 12016 launches (warm-up step + timed step) took 20 minutes; skip with SKIP_LAUNCH_LIST=1.
if [ -z "$SKIP_LAUNCH_LIST" ]; then
python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/plain_n65536.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 40000 --csv --log-file gpurun_out/launches_r1_n65536.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_launches_n65536.log 2>&1
fi
# 2. config 5 fused step: the fused Lbar + Adam kernel and the three GEMM shapes
python tools/configs_probe.py 5f > gpurun_out/plain_5f.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:tril_rank_adam -s 3 -c 1 -o gpurun_out/prof_tril_rank_adam_r1${SUFFIX} \
    python tools/configs_probe.py 5f > gpurun_out/ncu_5f_a.log 2>&1
if [ -z "$SKIP_GEMMS" ]; then
ncu --set full --clock-control none --import-source on -k regex:gemm_tc2_kernel -s 9 -c 3 -o gpurun_out/prof_linop_gemms_r1 \
    python tools/configs_probe.py 5f > gpurun_out/ncu_5f_b.log 2>&1
fi
# 3. config 4: the density family kernels at [32*4096, 784]
python tools/configs_probe.py 4 > gpurun_out/plain_4.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:density_ -s 6 -c 2 -o gpurun_out/prof_density_r1${SUFFIX} \
    python tools/configs_probe.py 4 > gpurun_out/ncu_4.log 2>&1
tail -n 3 gpurun_out/plain_5f.log; tail -n 1 gpurun_out/plain_4.log
