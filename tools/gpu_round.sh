#!/bin/bash
# One GPU trip that regenerates the round's evidence from the current tree: tests, smoke, the bench lines of every BASELINE
# config (the default line carries C2 / C4 pipelined / C5 under `other_configs`), micro-benchmarks.  Raw logs under
# gpurun_out/; copy into profiles/ with tools/collect_profiles.sh afterwards (here, no GPU needed).
mkdir -p gpurun_out
rm -f gpurun_out/parity_numbers.jsonl
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -5 | tee gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee gpurun_out/smoke.log
timeout 900 python bench.py --steps 3 --warmup 3 2>gpurun_out/bench.err | tail -1 > gpurun_out/bench_c3.json; cut -c1-300 gpurun_out/bench_c3.json
timeout 300 python bench.py --workload c2 --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/bench_c2.json; cut -c1-200 gpurun_out/bench_c2.json
COVO_FLOW_PERSISTENT=1 timeout 300 python bench.py --workload c2 --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/bench_c2_persistent.json; cut -c1-200 gpurun_out/bench_c2_persistent.json
timeout 600 python bench.py --workload c4 --steps 2 --warmup 3 2>/dev/null | tail -1 > gpurun_out/bench_c4.json; cut -c1-200 gpurun_out/bench_c4.json
timeout 600 python bench.py --workload c4p --steps 3 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/bench_c4p.json; cut -c1-200 gpurun_out/bench_c4p.json
timeout 600 python bench.py --workload c5 --steps 3 2>/dev/null | grep '^{' > gpurun_out/bench_c5.jsonl; wc -l gpurun_out/bench_c5.jsonl
timeout 200 python tools/gemm_bench.py 2>&1 | tee gpurun_out/gemm_bench.txt | tail -3
timeout 200 python tools/gemm_epi_bench.py 2>&1 | tee gpurun_out/gemm_epi_bench.txt | tail -3
timeout 100 python tools/gemm_trace.py 2>&1 | grep 'gemm trace' > gpurun_out/gemm_trace.txt; wc -l gpurun_out/gemm_trace.txt
timeout 100 python tools/gemm_narrow_bench.py > gpurun_out/gemm_narrow_bench.txt 2>&1; wc -l gpurun_out/gemm_narrow_bench.txt
timeout 100 tools/micro/attn_bench quick 2>&1 | tee gpurun_out/attn_bench.txt | tail -3
timeout 60 tools/micro/attn_trace 2>&1 | tee gpurun_out/attn_trace.txt | tail -3
COVO_FLOW_PERSISTENT=1 COVO_FLOW_TRACE=1 timeout 120 python bench.py --workload c2 --steps 1 --warmup 1 --no-cpu-baseline 2>&1 | grep "flow trace" > gpurun_out/flow_persistent_trace.txt; wc -l gpurun_out/flow_persistent_trace.txt
T2S_FMTS=bf16,fp32 timeout 300 python tools/t2s_bench.py comix 1500 2>&1 | grep "B=" | tee gpurun_out/t2s_bench.log
timeout 100 python tools/vocoder_bench.py 2>&1 | tee gpurun_out/vocoder_bench.txt
