#!/bin/bash
# Round-2 second GPU call (2 GPUs): the whole GPU suite after the tail-batch fix, the NCCL tests, A/B of the bulk-staged
# GroupNorm backward and of the overlapped bucket all-reduce.
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
run() { name=$1; t=$2; shift 2; echo "== $name"; timeout $t "$@" > gpurun_out/r2c2_$name.log 2>&1; rc=$?; echo "$name rc=$rc"; tail -${TAILN:-4} gpurun_out/r2c2_$name.log | cut -c1-400; return $rc; }
export CUDA_VISIBLE_DEVICES=0,1
TAILN=25 run pytest_gpu 1500 python -m pytest tests -m gpu -q -x
TAILN=15 run diag_default 300 python tools/diag_full_grad.py cifar10_b128
short="--steps 30 --warmup 8 --no-cpu --no-eager --no-secondary --sample-steps 100 --sustain-s 0"
CUDA_VISIBLE_DEVICES=0 TAILN=2 run bench1_bulk1 400 python bench.py $short
CUDA_VISIBLE_DEVICES=0 IGM_GN_BULK=0 TAILN=2 run bench1_bulk0 400 python bench.py $short
CUDA_VISIBLE_DEVICES=0 TAILN=2 run bench1_secondary 600 python bench.py --steps 10 --warmup 3 --no-cpu --no-eager --sample-steps 50 --sustain-s 0
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
TAILN=2 run bench2_overlap1 600 $TR bench.py --gpus 2 $short
IGM_DDP_OVERLAP=0 TAILN=2 run bench2_overlap0 600 $TR bench.py --gpus 2 $short
TAILN=2 run bench2_celeba_overlap1 600 $TR bench.py --gpus 2 --config celeba64 $short
IGM_DDP_OVERLAP=0 TAILN=2 run bench2_celeba_overlap0 600 $TR bench.py --gpus 2 --config celeba64 $short
TAILN=2 run bench2_vqvae 600 $TR bench.py --gpus 2 --config vqvae --steps 20 --warmup 5
CUDA_VISIBLE_DEVICES=0 TAILN=2 run bench1_vqvae 600 python bench.py --config vqvae --steps 20 --warmup 5
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/r2c2_bench*.log')):
    try:
        d = json.loads([l for l in open(f) if l.startswith('{')][-1])
        cls = {k: round(v['ms_per_step'], 3) for k, v in (d.get('roofline') or {}).get('classes', {}).items()}
        print(f.split('r2c2_')[1], 'value', round(d['value'], 2), 'ms', round(d['ms_per_step'], 3), 'e2e', round(d['e2e']['value'], 2),
              'samples/s', d.get('samples_per_sec_1000step'), cls,
              {k: (d[k].get('value') if 'value' in d[k] else {kk: vv.get('value') for kk, vv in d[k].items() if isinstance(vv, dict)}) for k in ('pixelcnn', 'vqvae') if k in d})
    except Exception as e:
        print(f, 'no json', e)
PY
echo done
