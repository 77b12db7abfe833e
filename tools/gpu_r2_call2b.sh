#!/bin/bash
# Round-2 GPU call 2b (2 GPUs): NCCL parity test, overlapped bucket all-reduce A/B at 2 ranks, VQ-VAE data parallel.
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
run() { name=$1; t=$2; shift 2; echo "== $name"; timeout $t "$@" > gpurun_out/r2c2_$name.log 2>&1; rc=$?; echo "$name rc=$rc"; tail -${TAILN:-4} gpurun_out/r2c2_$name.log | cut -c1-400; return $rc; }
TAILN=25 run pytest_nccl 600 python -m pytest tests/test_gpu_nccl.py -m gpu -q -x
cat gpurun_out/nccl_parity.txt
short="--steps 30 --warmup 8 --no-cpu --no-eager --no-secondary --sample-steps 100 --sustain-s 0"
N=${NGPU:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
TAILN=2 run bench${N}_overlap1 600 $TR bench.py --gpus $N $short
IGM_DDP_OVERLAP=0 TAILN=2 run bench${N}_overlap0 600 $TR bench.py --gpus $N $short
TAILN=2 run bench${N}_overlap1b 600 $TR bench.py --gpus $N $short
TAILN=2 run bench${N}_celeba_overlap1 600 $TR bench.py --gpus $N --config celeba64 $short
IGM_DDP_OVERLAP=0 TAILN=2 run bench${N}_celeba_overlap0 600 $TR bench.py --gpus $N --config celeba64 $short
TAILN=2 run bench${N}_vqvae 600 $TR bench.py --gpus $N --config vqvae --steps 20 --warmup 5
python tools/summarize_bench_logs.py gpurun_out/r2c2_bench${N}*.log
echo done
