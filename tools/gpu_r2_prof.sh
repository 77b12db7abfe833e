#!/bin/bash
# Round-2 profiling call (1 GPU): ncu launch list of one training + one sampler step, and --set full captures of the
# top kernels of each class.  Numbers printed by runs under ncu are never bench values.
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
P="python tools/profile_step.py"
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches.csv $P --what both > gpurun_out/r2_prof_launches.log 2>&1
echo "launch list rc=$?"
full() { name=$1; pat=$2; skip=$3; cnt=$4; timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:$pat -s $skip -c $cnt -f -o gpurun_out/r2_$name $P --what train > gpurun_out/r2_prof_$name.log 2>&1; echo "$name rc=$?"; }
full gn_bwd gn_bwd_bulk_kernel 0 3
full gn_apply gn_apply_fast_kernel 0 4
full conv_tc conv_tc_kernel 2 8
full wgrad_halo wgrad_halo_kernel 0 4
full attn_bwd linattn_bwd 0 2
full attn_fwd "linattn_(ctx|out)_kernel" 0 2
ls -la gpurun_out/*.ncu-rep
echo done
