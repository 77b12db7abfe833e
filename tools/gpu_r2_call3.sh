#!/bin/bash
# Round-2 GPU call 3 (1 GPU): tensor-core attention path: parity suite, then A/B against the CUDA-core kernels.
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
run() { name=$1; t=$2; shift 2; echo "== $name"; timeout $t "$@" > gpurun_out/r2c3_$name.log 2>&1; rc=$?; echo "$name rc=$rc"; tail -${TAILN:-4} gpurun_out/r2c3_$name.log | cut -c1-600; return $rc; }
TAILN=40 run pytest_parity 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "cifar10 or celeba64 or tiny"
TAILN=25 run pytest_gpu 1500 python -m pytest tests -m gpu -q -x
short="--steps 30 --warmup 8 --no-cpu --no-eager --no-secondary --sample-steps 100 --sustain-s 0"
TAILN=2 run bench_atc1 400 python bench.py $short
IGM_ATTN_TC=0 TAILN=2 run bench_atc0 400 python bench.py $short
TAILN=2 run bench_celeba_atc1 400 python bench.py --config celeba64 $short
IGM_ATTN_TC=0 TAILN=2 run bench_celeba_atc0 400 python bench.py --config celeba64 $short
python tools/summarize_bench_logs.py gpurun_out/r2c3_bench*.log
echo done
