#!/bin/bash
# Round-2 GPU call 4 (2 GPUs): suite + NCCL test (direct-into-.grad data-parallel path), 64-wide tile experiment,
# attention threshold, NCCL stream priority experiment.
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
run() { name=$1; t=$2; shift 2; echo "== $name"; timeout $t "$@" > gpurun_out/r2c4_$name.log 2>&1; rc=$?; echo "$name rc=$rc"; tail -${TAILN:-4} gpurun_out/r2c4_$name.log | cut -c1-500; return $rc; }
TAILN=25 run pytest_gpu 1500 python -m pytest tests -m gpu -q -x
cat gpurun_out/nccl_parity.txt | tail -2
export CUDA_VISIBLE_DEVICES=0
short="--steps 30 --warmup 8 --no-cpu --no-eager --no-secondary --sample-steps 200 --sustain-s 0"
TAILN=1 run bench_base 400 python bench.py $short
IGM_TC_BN=64 TAILN=1 run bench_bn64 400 python bench.py $short
TAILN=1 run bench_base2 400 python bench.py $short
IGM_TC_BN=64 TAILN=1 run bench_bn64b 400 python bench.py $short
TAILN=1 run bench_celeba 400 python bench.py --config celeba64 $short
IGM_TC_BN=64 TAILN=1 run bench_celeba_bn64 400 python bench.py --config celeba64 $short
IGM_ATTN_TC=0 TAILN=1 run bench_celeba_atc0 400 python bench.py --config celeba64 $short
export CUDA_VISIBLE_DEVICES=0,1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
TAILN=1 run bench2_base 600 $TR bench.py --gpus 2 $short
IGM_NCCL_PRIO=1 TAILN=1 run bench2_prio 600 $TR bench.py --gpus 2 $short
IGM_DDP_OVERLAP=0 TAILN=1 run bench2_overlap0 600 $TR bench.py --gpus 2 $short
python tools/summarize_bench_logs.py gpurun_out/r2c4_bench*.log
echo done
