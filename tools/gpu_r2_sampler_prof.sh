#!/bin/bash
# ncu launch list of ONE sampler step (B = 128, CIFAR-10 config): per-launch durations in launch order
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2s_sample_launches.csv python tools/profile_step.py --what sample > gpurun_out/r2s_prof.log 2>&1; echo "rc=$?"
python tools/summarize_ncu.py gpurun_out/r2s_sample_launches.csv | head -40
