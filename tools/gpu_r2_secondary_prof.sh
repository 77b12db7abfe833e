#!/bin/bash
# ncu --set full captures of the secondary paths' own kernels: the PixelCNN engine, the VQ lookup, the thin stride-2 layers
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on \
  -k regex:"pixelcnn_kernel|vq_forward_kernel|vq_backward_kernel|thin_in_s2|thin_out_s2|wgrad_thin_s2" -c 8 -f -o gpurun_out/r2_secondary \
  python tools/profile_secondary.py > gpurun_out/r2_prof_secondary.log 2>&1
echo "secondary rc=$?"
ls -la gpurun_out/r2_secondary.ncu-rep
