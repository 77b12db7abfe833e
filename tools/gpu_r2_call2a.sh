#!/bin/bash
# Round-2 GPU call 2a (1 GPU): the whole GPU suite after the tail-batch fix, A/B of the bulk-staged GroupNorm backward,
# the secondary legs with the wide-load PixelCNN GEMVs.
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
run() { name=$1; t=$2; shift 2; echo "== $name"; timeout $t "$@" > gpurun_out/r2c2_$name.log 2>&1; rc=$?; echo "$name rc=$rc"; tail -${TAILN:-4} gpurun_out/r2c2_$name.log | cut -c1-400; return $rc; }
TAILN=25 run pytest_gpu 1500 python -m pytest tests -m gpu -q -x
TAILN=15 run diag_default 300 python tools/diag_full_grad.py cifar10_b128
short="--steps 30 --warmup 8 --no-cpu --no-eager --no-secondary --sample-steps 100 --sustain-s 0"
TAILN=2 run bench1_bulk1 400 python bench.py $short
IGM_GN_BULK=0 TAILN=2 run bench1_bulk0 400 python bench.py $short
TAILN=2 run bench1_bulk1b 400 python bench.py $short
TAILN=2 run bench1_secondary 600 python bench.py --steps 10 --warmup 3 --no-cpu --no-eager --sample-steps 50 --sustain-s 0
TAILN=2 run bench1_vqvae 600 python bench.py --config vqvae --steps 20 --warmup 5
python tools/summarize_bench_logs.py gpurun_out/r2c2_bench1*.log
echo done
