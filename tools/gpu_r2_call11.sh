#!/bin/bash
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
run() { name=$1; t=$2; shift 2; echo "== $name"; timeout $t "$@" > gpurun_out/r2c11_$name.log 2>&1; rc=$?; echo "$name rc=$rc"; tail -${TAILN:-4} gpurun_out/r2c11_$name.log | cut -c1-400; return $rc; }
TAILN=25 run pytest_gpu 1500 python -m pytest tests -m gpu -q -x
short="--steps 30 --warmup 8 --no-cpu --no-eager --no-secondary --sample-steps 200 --sustain-s 0"
TAILN=1 run bench 600 python bench.py $short
TAILN=1 run bench_b 600 python bench.py $short
TAILN=1 run bench_celeba 600 python bench.py --config celeba64 $short
python tools/summarize_bench_logs.py gpurun_out/r2c11_bench*.log
echo done
