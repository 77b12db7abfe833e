#!/bin/bash
# Round-2 GPU call 6 (1 GPU): generic conv operator on the tensor cores incl. ConvTranspose k4 s2; whole suite; secondary bench.
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
run() { name=$1; t=$2; shift 2; echo "== $name"; timeout $t "$@" > gpurun_out/r2c6_$name.log 2>&1; rc=$?; echo "$name rc=$rc"; tail -${TAILN:-4} gpurun_out/r2c6_$name.log | cut -c1-700; return $rc; }
TAILN=40 run pytest_ops 900 python -m pytest tests/test_ops.py tests/test_vqvae.py tests/test_vq.py tests/test_pixelcnn.py -m gpu -q
TAILN=25 run pytest_gpu 1500 python -m pytest tests -m gpu -q -x
TAILN=2 run bench_vqvae 600 python bench.py --config vqvae --steps 20 --warmup 5
TAILN=2 run bench_secondary 600 python bench.py --steps 10 --warmup 3 --no-cpu --no-eager --sample-steps 50 --sustain-s 0
python tools/summarize_bench_logs.py gpurun_out/r2c6_bench*.log
echo done
