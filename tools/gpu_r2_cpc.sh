#!/bin/bash
# chunks-per-CTA of the attention statistics / context kernels: suite, then IGM_ATTN_CPC=1 (one chunk per CTA) against the
# default rule, alternating inside ONE box (box-to-box noise is +-0.3 %)
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r2cpc_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r2cpc_pytest.log
short="--steps 40 --warmup 8 --no-cpu --no-eager --no-secondary --sample-steps 100 --sustain-s 0"
for r in a b; do
  IGM_ATTN_CPC=1 timeout 600 python bench.py $short > gpurun_out/r2cpc_bench_cpc1_$r.log 2>&1
  timeout 600 python bench.py $short > gpurun_out/r2cpc_bench_auto_$r.log 2>&1
done
python tools/summarize_bench_logs.py gpurun_out/r2cpc_bench_*.log
