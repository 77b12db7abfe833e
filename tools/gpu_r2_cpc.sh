#!/bin/bash
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
short="--steps 20 --warmup 6 --no-cpu --no-eager --no-secondary --sample-steps 100 --sustain-s 0"
for c in 1 2 4 8; do IGM_ATTN_CPC=$c timeout 600 python bench.py $short > gpurun_out/r2cpc_bench_$c.log 2>&1; done
python tools/summarize_bench_logs.py gpurun_out/r2cpc_bench_*.log
