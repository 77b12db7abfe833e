#!/bin/bash
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
run() { name=$1; t=$2; shift 2; echo "== $name"; timeout $t "$@" > gpurun_out/r2c7_$name.log 2>&1; rc=$?; echo "$name rc=$rc"; tail -${TAILN:-4} gpurun_out/r2c7_$name.log | cut -c1-700; return $rc; }
TAILN=30 run pytest_vq 900 python -m pytest tests/test_vq.py tests/test_vqvae.py -m gpu -q
python - <<'PY'
import sys, json
sys.path.insert(0, '.')
import torch, bench, bench_secondary
dev = torch.device('cuda', 0)
out = bench_secondary.vq_leg(dev, bench.peaks(), cpu=False)
print(json.dumps(out['vq_lookup']))
PY
echo done
