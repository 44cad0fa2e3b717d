#!/bin/bash
# Profiling trip for the text-to-semantic decode kernel and the C4 pipeline: ncu launch list of the C4 bench command + one
# --set full capture of t2s_decode_kernel (200 positions, B = 8).  Logs under gpurun_out/.
mkdir -p gpurun_out
export COVO_NO_GRAPH=1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 700 --csv --log-file gpurun_out/launches_c4.csv \
    python bench.py --workload c4 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/bench_c4_under_ncu.log 2>&1
unset COVO_NO_GRAPH
T2S_FMTS=bf16 T2S_B=8 timeout 900 ncu --set full --clock-control none --import-source on -k regex:t2s_decode -s 1 -c 1 \
    -o gpurun_out/t2s_decode python tools/t2s_bench.py comix 200 > gpurun_out/t2s_under_ncu.log 2>&1
python tools/summarize_ncu.py launches gpurun_out/launches_c4.csv gpurun_out/launches_c4.md > /dev/null
python tools/summarize_ncu.py full gpurun_out/t2s_decode.ncu-rep gpurun_out/t2s_decode_ncu.md > /dev/null
rm -f gpurun_out/t2s_decode.ncu-rep
ls -la gpurun_out
