#!/bin/bash
# A/B timing of one environment switch: tools/ab_bench.sh VAR  ->  bench with VAR=0 and VAR unset (no CPU leg, short sampler)
mkdir -p gpurun_out
VAR=$1; TAG=${2:-ab}
for v in 0 1; do
  if [ $v == 0 ]; then export $VAR=0; else unset $VAR; fi
  timeout 600 python bench.py --steps 30 --warmup 8 --no-cpu --sample-steps 100 > gpurun_out/bench_${TAG}$v.json 2> gpurun_out/bench_${TAG}$v.err
  echo "$VAR=$v rc=$?"; python - <<PY
import json
d=json.load(open("gpurun_out/bench_${TAG}$v.json"))
print("steps/s %.1f  ms %.3f  e2e %.1f  samples/s %.1f" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["samples_per_sec_1000step"]))
PY
done
