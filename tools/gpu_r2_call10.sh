#!/bin/bash
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
run() { name=$1; t=$2; shift 2; echo "== $name"; timeout $t "$@" > gpurun_out/r2c10_$name.log 2>&1; rc=$?; echo "$name rc=$rc"; tail -${TAILN:-4} gpurun_out/r2c10_$name.log | cut -c1-500; return $rc; }
TAILN=30 run pytest_sampler 900 python -m pytest tests/test_gpu_parity.py tests/test_zz_gpu_full_size.py -m gpu -q -x -k "sampler or chain or smoke"
TAILN=25 run pytest_gpu 1500 python -m pytest tests -m gpu -q -x
short="--steps 20 --warmup 5 --no-cpu --no-eager --no-secondary --sustain-s 0"
TAILN=1 run bench 600 python bench.py $short
IGM_GN_EPI=0 TAILN=1 run bench_epi0 600 python bench.py $short
TAILN=1 run bench_celeba 600 python bench.py --config celeba64 $short
IGM_GN_EPI=0 TAILN=1 run bench_celeba_epi0 600 python bench.py --config celeba64 $short
python tools/summarize_bench_logs.py gpurun_out/r2c10_bench*.log
echo done
