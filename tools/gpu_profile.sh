#!/bin/bash
# Profiling trip: launch list of the bench command + full captures of the two dominant kernels (and of the vocoder's
# dominant kernels at the bench shape).  Raw artefacts under gpurun_out/; tools/summarize_ncu.py turns them into
# profiles/r02_*.md here (no GPU needed).  Numbers printed under ncu are never bench values.
mkdir -p gpurun_out
export COVO_NO_GRAPH=1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 420 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-other-configs > gpurun_out/bench_under_ncu.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tc -s 40 -c 8 -o gpurun_out/gemm_c3 -f \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-other-configs > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attention_tc -s 4 -c 1 -o gpurun_out/attn_c3 -f \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-other-configs > /dev/null 2>&1
unset COVO_NO_GRAPH
# vocoder at the bench shape [8, 80, 1500]: DRAM bytes of every launch of one forward (metrics only: fast)
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
    --log-file gpurun_out/voc_dram.csv python tools/vocoder_bench.py one > gpurun_out/voc_under_ncu.log 2>&1
# the fused last vocoder stage at the bench shape (51 tiles per SM, not the wave-tail case B = 1, T = 256)
timeout 600 ncu --set full --clock-control none --import-source on -k regex:hifigan_fused -s 2 -c 1 -o gpurun_out/hifigan_fused -f \
    python tools/vocoder_bench.py one > /dev/null 2>&1
timeout 300 python tools/attn_one.py | tee gpurun_out/attn_one.log
# summarise on the box and drop the reports (gpurun copies back at most 64 MiB); the GEMM report is kept for the source page
for k in gemm_c3 attn_c3 hifigan_fused; do
  [ -s gpurun_out/$k.ncu-rep ] && python tools/summarize_ncu.py full gpurun_out/$k.ncu-rep gpurun_out/${k}_ncu.md > /dev/null
done
python tools/summarize_ncu.py launches gpurun_out/launches.csv gpurun_out/launches_c3.md > /dev/null
python tools/summarize_ncu.py dram gpurun_out/voc_dram.csv gpurun_out/vocoder_dram_ncu.md > /dev/null
rm -f gpurun_out/attn_c3.ncu-rep gpurun_out/hifigan_fused.ncu-rep
ls -la gpurun_out | tail -20
