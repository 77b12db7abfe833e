#!/bin/bash
# within-one-box A/B of an environment switch: AB_ENV="NAME=value" against the default, alternating, both DDPM configs
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
T=${TAG:-r2ab}
short="--steps 40 --warmup 8 --no-cpu --no-eager --no-secondary --sample-steps 100 --sustain-s 0"
if [ "${SUITE:-0}" == "1" ]; then env $AB_ENV timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/${T}_pytest.log 2>&1; echo "pytest (with $AB_ENV) rc=$?"; tail -2 gpurun_out/${T}_pytest.log; fi
for r in a b; do
  env $AB_ENV timeout 600 python bench.py $short > gpurun_out/${T}_bench_on_$r.log 2>&1
  timeout 600 python bench.py $short > gpurun_out/${T}_bench_off_$r.log 2>&1
done
env $AB_ENV timeout 600 python bench.py --config celeba64 $short > gpurun_out/${T}_bench_celeba_on.log 2>&1
timeout 600 python bench.py --config celeba64 $short > gpurun_out/${T}_bench_celeba_off.log 2>&1
python tools/summarize_bench_logs.py gpurun_out/${T}_bench_*.log
