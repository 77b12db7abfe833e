#!/bin/bash
# Round-2 scaling call (N GPUs, NGPU env): overlapped bucket all-reduce A/B for the two DDPM configs, VQ-VAE data parallel.
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
N=${NGPU:-8}
run() { name=$1; t=$2; shift 2; echo "== $name"; timeout $t "$@" > gpurun_out/r2sc_$name.log 2>&1; rc=$?; echo "$name rc=$rc"; tail -${TAILN:-2} gpurun_out/r2sc_$name.log | cut -c1-300; return $rc; }
short="--steps 30 --warmup 8 --no-cpu --no-eager --no-secondary --sample-steps 100 --sustain-s 0"
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
run bench${N}_overlap1 600 $TR bench.py --gpus $N $short
IGM_DDP_OVERLAP=0 run bench${N}_overlap0 600 $TR bench.py --gpus $N $short
run bench${N}_overlap1b 600 $TR bench.py --gpus $N $short
IGM_DDP_OVERLAP=0 run bench${N}_overlap0b 600 $TR bench.py --gpus $N $short
run bench${N}_celeba_overlap1 600 $TR bench.py --gpus $N --config celeba64 $short
IGM_DDP_OVERLAP=0 run bench${N}_celeba_overlap0 600 $TR bench.py --gpus $N --config celeba64 $short
run bench${N}_vqvae 600 $TR bench.py --gpus $N --config vqvae --steps 20 --warmup 5
python tools/summarize_bench_logs.py gpurun_out/r2sc_bench${N}*.log
echo done
