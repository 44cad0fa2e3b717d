#!/bin/bash
# 8-GPU evidence (BASELINE configs[3] and the 8-GPU arm of configs[4]): run with `gpurun --gpus 8 -- bash tools/gpu_round_8gpu.sh`.
# Raw lines under gpurun_out/; copy to profiles/r02_bench_c4p_8gpu.json / r02_bench_c5_vocoder_sweep_8gpu.jsonl afterwards.
mkdir -p gpurun_out
N=${1:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
timeout 600 $TR bench.py --gpus $N --workload c4p --steps 3 --warmup 3 --no-cpu-baseline 2>gpurun_out/bench_c4p_${N}gpu.err | grep '^{' | tail -1 > gpurun_out/bench_c4p_${N}gpu.json
cut -c1-260 gpurun_out/bench_c4p_${N}gpu.json
timeout 600 $TR bench.py --gpus $N --workload c5 --steps 3 2>gpurun_out/bench_c5_${N}gpu.err | grep '^{' > gpurun_out/bench_c5_${N}gpu.jsonl
wc -l gpurun_out/bench_c5_${N}gpu.jsonl; tail -2 gpurun_out/bench_c5_${N}gpu.jsonl | cut -c1-260
timeout 300 $TR bench.py --gpus $N --steps 3 --warmup 3 --no-cpu-baseline --no-other-configs 2>gpurun_out/bench_c3_${N}gpu.err | grep '^{' | tail -1 > gpurun_out/bench_c3_${N}gpu.json
cut -c1-260 gpurun_out/bench_c3_${N}gpu.json
