#!/bin/bash
# two short benches per DDPM config (A/B of a build against the previous call's numbers)
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
T=${TAG:-r2q}
run() { name=$1; t=$2; shift 2; timeout $t "$@" > gpurun_out/${T}_$name.log 2>&1; rc=$?; echo "$name rc=$rc"; return $rc; }
short="--steps 30 --warmup 8 --no-cpu --no-eager --no-secondary --sample-steps 100 --sustain-s 0"
run bench_a 600 python bench.py $short
run bench_b 600 python bench.py $short
run bench_celeba_a 600 python bench.py --config celeba64 $short
python tools/summarize_bench_logs.py gpurun_out/${T}_bench*.log
