#!/bin/bash
# Copy the raw artefacts of tools/gpu_round.sh / tools/gpu_profile.sh (gpurun_out/) into profiles/ under round-2 names.
R=${1:-r02}
cp gpurun_out/bench_c3.json profiles/${R}_bench_c3.json
cp gpurun_out/bench_c2.json profiles/${R}_bench_c2.json
cp gpurun_out/bench_c2_persistent.json profiles/${R}_bench_c2_persistent_kernel.json
cp gpurun_out/bench_c4.json profiles/${R}_bench_c4_pipeline.json
cp gpurun_out/bench_c4p.json profiles/${R}_bench_c4p_pipelined.json
cp gpurun_out/bench_c5.jsonl profiles/${R}_bench_c5_vocoder_sweep.jsonl
cp gpurun_out/gemm_bench.txt profiles/${R}_gemm_microbench.txt
cp gpurun_out/gemm_epi_bench.txt profiles/${R}_gemm_epilogue_cost.txt
[ -s gpurun_out/gemm_trace.txt ] && cp gpurun_out/gemm_trace.txt profiles/${R}_gemm_trace.txt
[ -s gpurun_out/gemm_narrow_bench.txt ] && cp gpurun_out/gemm_narrow_bench.txt profiles/${R}_gemm_narrow_tiles.txt
cp gpurun_out/attn_bench.txt profiles/${R}_attention_bench.txt
cp gpurun_out/attn_trace.txt profiles/${R}_attention_trace.txt
cp gpurun_out/flow_persistent_trace.txt profiles/${R}_flow_persistent_trace.txt
cp gpurun_out/t2s_bench.log profiles/${R}_t2s_bench.txt
cp gpurun_out/vocoder_bench.txt profiles/${R}_vocoder_bench.txt
cp gpurun_out/parity_numbers.jsonl profiles/${R}_parity_numbers.jsonl
cp gpurun_out/pytest_gpu.log profiles/${R}_pytest_gpu.log
python tools/sass_tally.py > profiles/${R}_sass_tally.md
ls profiles | grep ${R}
# 8-GPU evidence (tools/gpu_round_8gpu.sh), when present
for f in bench_c4p_8gpu.json bench_c3_8gpu.json; do [ -s gpurun_out/$f ] && cp gpurun_out/$f profiles/${R}_$f; done
[ -s gpurun_out/bench_c5_8gpu.jsonl ] && cp gpurun_out/bench_c5_8gpu.jsonl profiles/${R}_bench_c5_vocoder_sweep_8gpu.jsonl
