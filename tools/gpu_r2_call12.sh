#!/bin/bash
# A/B of the k | v-only to_qkv conv on the sampler's per-image-matrix attention path (IGM_ATTN_KV=0 restores q | k | v)
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
run() { name=$1; t=$2; shift 2; echo "== $name"; timeout $t "$@" > gpurun_out/r2c12_$name.log 2>&1; rc=$?; echo "$name rc=$rc"; tail -${TAILN:-4} gpurun_out/r2c12_$name.log | cut -c1-400; return $rc; }
TAILN=25 run pytest_gpu 1500 python -m pytest tests -m gpu -q -x
short="--steps 10 --warmup 4 --no-cpu --no-eager --no-secondary --sample-steps 1000 --sustain-s 0"
TAILN=1 run bench_kv1 600 python bench.py $short
IGM_ATTN_KV=0 TAILN=1 run bench_kv0 600 python bench.py $short
TAILN=1 run bench_kv1b 600 python bench.py $short
TAILN=1 run bench_celeba_kv1 600 python bench.py --config celeba64 $short
IGM_ATTN_KV=0 TAILN=1 run bench_celeba_kv0 600 python bench.py --config celeba64 $short
python tools/summarize_bench_logs.py gpurun_out/r2c12_bench*.log
echo done
