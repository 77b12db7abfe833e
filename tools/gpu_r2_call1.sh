#!/bin/bash
# Round-2 first GPU call: root-cause of the red full-size gradient test, the whole GPU suite, the bench line with the
# eager-GPU comparator and the secondary metrics, then the first execution of the two bring-up conv kernels.
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r2_gpu.txt 2>&1
run() { name=$1; t=$2; shift 2; echo "== $name"; timeout $t "$@" > gpurun_out/r2c1_$name.log 2>&1; rc=$?; echo "$name rc=$rc"; tail -${TAILN:-4} gpurun_out/r2c1_$name.log; return $rc; }

TAILN=30 run diag_default 300 python tools/diag_full_grad.py cifar10_b128
IGM_WGRAD_STREAM=0 TAILN=8 run diag_nostream 300 python tools/diag_full_grad.py cifar10_b128
IGM_WGRAD_HALO=0 TAILN=8 run diag_nohalo 300 python tools/diag_full_grad.py cifar10_b128
IGM_PDL=0 TAILN=8 run diag_nopdl 300 python tools/diag_full_grad.py cifar10_b128
IGM_GN_FUSED=0 TAILN=8 run diag_nogn 300 python tools/diag_full_grad.py cifar10_b128
IGM_CONV_ENGINE=0 TAILN=8 run diag_simt 400 python tools/diag_full_grad.py cifar10_b128
TAILN=20 run diag_celeba 400 python tools/diag_full_grad.py celeba64_b32

TAILN=15 run pytest_gpu 1500 python -m pytest tests -m gpu -q -x
TAILN=3 run bench 900 python bench.py
tail -c 6000 gpurun_out/r2c1_bench.log
TAILN=3 run bench_celeba 600 python bench.py --config celeba64 --no-secondary
TAILN=3 run bench_ref 300 python bench.py --impl reference --steps 3 --warmup 1

# first execution of the bring-up kernels, one small shape each, short timeouts (a protocol bug hangs the kernel)
IGM_TEST_CONV_PAIR=1 TAILN=12 run pair_first 120 python -m pytest tests/test_gpu_conv_tc.py -m gpu -x -q -k "conv_pair_forward and shape0"; pair=$?
IGM_TEST_CONV_HALO=1 TAILN=12 run halo_first 120 python -m pytest tests/test_gpu_conv_tc.py -m gpu -x -q -k "conv_halo_forward and shape0"; halo=$?
if [ $pair == 0 ]; then IGM_TEST_CONV_PAIR=1 TAILN=12 run pair_kernel 300 python -m pytest tests/test_gpu_conv_tc.py -m gpu -x -q -k "conv_pair_forward or conv_pair_dgrad"; pair=$?; fi
if [ $halo == 0 ]; then IGM_TEST_CONV_HALO=1 TAILN=12 run halo_kernel 300 python -m pytest tests/test_gpu_conv_tc.py -m gpu -x -q -k "conv_halo_forward or conv_halo_dgrad"; halo=$?; fi
flags=""; [ $halo == 0 ] && flags="$flags --halo"; [ $pair == 0 ] && flags="$flags --pair"
TAILN=60 run layer_bench 300 python tools/conv_layer_bench.py $flags
nvidia-smi > gpurun_out/r2c1_smi_end.txt 2>&1
echo done
