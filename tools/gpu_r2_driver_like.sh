#!/bin/bash
# The driver's own scaling invocation (both arms, default legs) on N GPUs: python -m torch.distributed.run ... bench.py --gpus N
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
N=${NGPU:-2}
T=${TAG:-r2drv}
run() { name=$1; t=$2; shift 2; echo "== $name"; s=$(date +%s); timeout $t "$@" > gpurun_out/${T}_$name.log 2>&1; rc=$?; echo "$name rc=$rc wall=$(( $(date +%s) - s ))s"; tail -1 gpurun_out/${T}_$name.log | cut -c1-400; return $rc; }
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"
run ref$N 900 $TR bench.py --impl reference --gpus $N --steps 20 --warmup 5
run ours$N 900 $TR bench.py --gpus $N --steps 20 --warmup 5
if [ "${CELEBA:-0}" == "1" ]; then run celeba$N 900 $TR bench.py --gpus $N --config celeba64 --steps 20 --warmup 5 --no-secondary; fi
python tools/summarize_bench_logs.py gpurun_out/${T}_ours$N.log
echo done
