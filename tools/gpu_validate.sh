#!/bin/bash
# suite + short benches (both DDPM configs), two runs each
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
T=${TAG:-r2v}
run() { name=$1; t=$2; shift 2; echo "== $name"; timeout $t "$@" > gpurun_out/${T}_$name.log 2>&1; rc=$?; echo "$name rc=$rc"; tail -${TAILN:-4} gpurun_out/${T}_$name.log | cut -c1-400; return $rc; }
TAILN=12 run pytest_gpu 1500 python -m pytest tests -m gpu -q -x
short="--steps 30 --warmup 8 --no-cpu --no-eager --no-secondary --sample-steps 100 --sustain-s 0"
TAILN=1 run bench_a 600 python bench.py $short
TAILN=1 run bench_b 600 python bench.py $short
TAILN=1 run bench_celeba_a 600 python bench.py --config celeba64 $short
TAILN=1 run bench_celeba_b 600 python bench.py --config celeba64 $short
python tools/summarize_bench_logs.py gpurun_out/${T}_bench*.log
echo done
