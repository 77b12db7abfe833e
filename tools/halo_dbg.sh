#!/bin/bash
# times the halo wgrad kernel under the bring-up switches (ncu launch list, wgrad_halo only)
mkdir -p gpurun_out
for d in ${HALO_DBG_LIST:-0 1 2 4 3}; do
  IGM_HALO_DEBUG=$d timeout 300 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv -k regex:wgrad_halo_kernel --log-file gpurun_out/halo_dbg_$d.csv python tools/profile_step.py --what train > /dev/null 2>&1
  echo "dbg=$d"; grep wgrad_halo gpurun_out/halo_dbg_$d.csv | awk -F'","' '{print $NF}' | tr -d '"' | head -8 | tr '\n' ' '; echo
done
