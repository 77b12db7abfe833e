#!/bin/bash
# persistent GroupNorm backward: forced-path parity, full suite, A/B against the per-sample cluster kernel
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
run() { name=$1; t=$2; shift 2; echo "== $name"; timeout $t "$@" > gpurun_out/r2c13_$name.log 2>&1; rc=$?; echo "$name rc=$rc"; tail -${TAILN:-4} gpurun_out/r2c13_$name.log | cut -c1-400; return $rc; }
TAILN=30 run gn_forced 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "persistent_groupnorm"
TAILN=12 run pytest_gpu 1500 python -m pytest tests -m gpu -q -x
short="--steps 30 --warmup 8 --no-cpu --no-eager --no-secondary --sample-steps 100 --sustain-s 0"
TAILN=1 run bench_p1 600 python bench.py $short
IGM_GN_PERSIST=0 TAILN=1 run bench_p0 600 python bench.py $short
TAILN=1 run bench_p1b 600 python bench.py $short
TAILN=1 run bench_celeba_p1 600 python bench.py --config celeba64 $short
IGM_GN_PERSIST=0 TAILN=1 run bench_celeba_p0 600 python bench.py --config celeba64 $short
python tools/summarize_bench_logs.py gpurun_out/r2c13_bench*.log
echo done
