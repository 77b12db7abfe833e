#!/bin/bash
# bench line + ncu launch list (+ optional full capture of the top kernel); logs under gpurun_out/
mkdir -p gpurun_out
TAG=${1:-r1}
echo "== bench" ; timeout 1200 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err ; echo "bench rc=$?" ; tail -3 gpurun_out/bench_${TAG}.err ; cat gpurun_out/bench_${TAG}.json
echo "== ncu launch list" ; timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_${TAG}.csv python tools/profile_step.py --what both > gpurun_out/ncu_launch_${TAG}.log 2>&1 ; echo "ncu rc=$?" ; tail -3 gpurun_out/ncu_launch_${TAG}.log
if [ -n "$2" ]; then
  echo "== ncu full: $2" ; timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:$2 -c 3 -f -o gpurun_out/prof_${TAG} python tools/profile_step.py --what train > gpurun_out/ncu_full_${TAG}.log 2>&1 ; echo "ncu full rc=$?" ; tail -3 gpurun_out/ncu_full_${TAG}.log
fi
