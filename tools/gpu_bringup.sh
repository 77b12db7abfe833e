#!/bin/bash
# One gpurun call that takes the two bring-up conv kernels (conv_halo.cu, conv_tc2.cu) from "never run" to measured:
#   /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash tools/gpu_bringup.sh'
# Every step runs under its own short timeout (a barrier-protocol bug hangs the kernel; the timeout kills the process
# long before gpurun's limit would make it a strike) and logs under gpurun_out/.  Steps are ordered cheapest-first and
# a failing kernel-level step skips the whole-net steps of that engine.
mkdir -p gpurun_out
step() { name=$1; t=$2; shift 2; echo "== $name"; timeout $t "$@" > gpurun_out/bringup_$name.log 2>&1; rc=$?; echo "$name rc=$rc"; tail -6 gpurun_out/bringup_$name.log; return $rc; }

# 1. smallest shapes first, one test at a time: the first launch of each kernel
IGM_TEST_CONV_HALO=1 step halo_first 180 python -m pytest tests/test_gpu_conv_tc.py -m gpu -x -q -k "conv_halo_forward and shape0"; halo=$?
IGM_TEST_CONV_PAIR=1 step pair_first 180 python -m pytest tests/test_gpu_conv_tc.py -m gpu -x -q -k "conv_pair_forward and shape0"; pair=$?

# 2. all kernel-level shapes
if [ $halo == 0 ]; then IGM_TEST_CONV_HALO=1 step halo_kernel 400 python -m pytest tests/test_gpu_conv_tc.py -m gpu -x -q -k "conv_halo_forward or conv_halo_dgrad"; halo=$?; fi
if [ $pair == 0 ]; then IGM_TEST_CONV_PAIR=1 step pair_kernel 400 python -m pytest tests/test_gpu_conv_tc.py -m gpu -x -q -k "conv_pair_forward or conv_pair_dgrad"; pair=$?; fi

# 3. per-layer timing of the engines that passed, side by side with the per-tap engine
flags=""; [ $halo == 0 ] && flags="$flags --halo"; [ $pair == 0 ] && flags="$flags --pair"
step layer_bench 300 python tools/conv_layer_bench.py $flags
step layer_bench_celeba 300 python tools/conv_layer_bench.py $flags --celeba

# 4. whole-net parity + A/B bench with the switch on
if [ $halo == 0 ]; then
  IGM_CONV_HALO=1 step halo_parity 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q
  IGM_CONV_HALO=1 step halo_bench 400 python bench.py --steps 30 --warmup 8 --no-cpu --sample-steps 100
fi
if [ $pair == 0 ]; then
  IGM_CONV_PAIR=1 step pair_parity 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q
  IGM_CONV_PAIR=1 step pair_bench 400 python bench.py --steps 30 --warmup 8 --no-cpu --sample-steps 100
fi
step base_bench 400 python bench.py --steps 30 --warmup 8 --no-cpu --sample-steps 100
