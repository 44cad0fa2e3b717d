#!/bin/bash
# Profiling trip: launch list of the bench command + full captures of the two dominant kernels.
mkdir -p gpurun_out
export COVO_NO_GRAPH=1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 420 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tc -s 40 -c 8 -o gpurun_out/gemm_c3 \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attention_tc -s 4 -c 1 -o gpurun_out/attn_c3 \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > /dev/null 2>&1
unset COVO_NO_GRAPH
timeout 300 python tools/attn_one.py | tee gpurun_out/attn_one.log
ls -la gpurun_out
