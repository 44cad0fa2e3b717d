#!/bin/bash
# One GPU trip: tests, smoke, the bench lines of every workload.  Logs under gpurun_out/.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -15 | tee gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee gpurun_out/smoke.log
timeout 600 python bench.py --steps 2 --warmup 3 2>&1 | tail -1 | tee gpurun_out/bench.log | cut -c1-400
timeout 300 python bench.py --workload c2 --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_c2.log | cut -c1-300
timeout 600 python bench.py --workload c4 --steps 2 --warmup 3 2>&1 | tail -1 | tee gpurun_out/bench_c4.log | cut -c1-300
timeout 600 python bench.py --workload c4p --steps 3 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_c4p.log | cut -c1-300
T2S_FMTS=bf16,fp32 timeout 300 python tools/t2s_bench.py comix 1500 2>&1 | grep "B=" | tee gpurun_out/t2s_bench.log
T2S_FMTS=bf16 timeout 300 python tools/t2s_bench.py cosingle 500 2>&1 | grep "B=" | tee -a gpurun_out/t2s_bench.log
timeout 600 python bench.py --workload c5 --steps 3 2>&1 | grep '^{' | tee gpurun_out/bench_c5.jsonl | cut -c1-200
