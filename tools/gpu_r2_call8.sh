#!/bin/bash
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
run() { name=$1; t=$2; shift 2; echo "== $name"; timeout $t "$@" > gpurun_out/r2c8_$name.log 2>&1; rc=$?; echo "$name rc=$rc"; tail -${TAILN:-4} gpurun_out/r2c8_$name.log | cut -c1-700; return $rc; }
TAILN=25 run pytest_gpu 1500 python -m pytest tests -m gpu -q -x
short="--steps 30 --warmup 8 --no-cpu --no-eager --sample-steps 200 --sustain-s 0"
TAILN=1 run bench 600 python bench.py $short
TAILN=1 run bench_celeba 600 python bench.py --config celeba64 --no-secondary $short
python tools/summarize_bench_logs.py gpurun_out/r2c8_bench*.log
python - <<'PY'
import json
d = json.loads([l for l in open('gpurun_out/r2c8_bench.log') if l.startswith('{')][-1])
print('pixelcnn', d['pixelcnn']['value'], 'vq c_abi', d['vqvae']['vq_lookup'].get('c_abi'), 'api', d['vqvae']['vq_lookup']['value'], 'roof', d['vqvae']['vq_lookup']['roofline'])
PY
echo done
