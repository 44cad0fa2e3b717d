#!/bin/bash
# One GPU trip: tests, bench, launch list.  Logs under gpurun_out/.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -15 | tee gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 2 --warmup 3 2>&1 | tail -3 | tee gpurun_out/bench.log
timeout 300 python bench.py --workload c2 --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_c2.log
