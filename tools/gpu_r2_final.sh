#!/bin/bash
# Round-2 final measurement call (1 GPU): suite, smoke, the default bench line and its reference arm, the other configs,
# ncu launch list + DRAM traffic of the dominant kernel.  Numbers printed under ncu are never bench values.
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
T=${TAG:-r2f}
run() { name=$1; t=$2; shift 2; echo "== $name"; timeout $t "$@" > gpurun_out/${T}_$name.log 2>&1; rc=$?; echo "$name rc=$rc"; tail -${TAILN:-3} gpurun_out/${T}_$name.log | cut -c1-400; return $rc; }
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv > gpurun_out/${T}_gpu.txt 2>&1
TAILN=6 run pytest_gpu 1500 python -m pytest tests -m gpu -q -x
TAILN=2 run smoke 300 python __graft_entry__.py smoke
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap --format=csv -lms 200 > gpurun_out/${T}_clocks.csv 2>&1 &
SMI=$!
TAILN=1 run bench 900 python bench.py
kill $SMI
TAILN=1 run bench_reference 600 python bench.py --impl reference
TAILN=1 run bench_celeba 900 python bench.py --config celeba64 --no-secondary
TAILN=1 run bench_vqvae 600 python bench.py --config vqvae
python tools/summarize_bench_logs.py gpurun_out/${T}_bench.log gpurun_out/${T}_bench_celeba.log gpurun_out/${T}_bench_vqvae.log
if [ "${NCU:-1}" == "1" ]; then
P="python tools/profile_step.py"
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${T}_launches.csv $P --what both > gpurun_out/${T}_prof_launches.log 2>&1; echo "launch list rc=$?"
timeout 900 ncu --profile-from-start off --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:conv_tc_kernel --csv --log-file gpurun_out/${T}_conv_dram.csv $P --what train > gpurun_out/${T}_prof_conv_dram.log 2>&1; echo "conv dram rc=$?"
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"linattn_(wt|p|dk)_kernel|linattn_bwd_dctx_kernel|linattn_ctx_kernel" -c 8 -f -o gpurun_out/${T}_attn_tc $P --what train > gpurun_out/${T}_prof_attn.log 2>&1; echo "attn rc=$?"
fi
echo done
