#!/bin/bash
# First-line GPU check: smoke, memcheck on the tiny config, full GPU parity suite.
# Everything is logged under gpurun_out/ (the only directory gpurun brings back).
mkdir -p gpurun_out
rm -f gpurun_out/parity_report.txt
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
echo "== smoke" ; timeout 600 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1 ; echo "smoke rc=$?" ; tail -5 gpurun_out/smoke.log
if [ "$1" == "sanitize" ]; then
  echo "== memcheck" ; timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python __graft_entry__.py smoke > gpurun_out/memcheck.log 2>&1 ; echo "memcheck rc=$?" ; tail -15 gpurun_out/memcheck.log
fi
echo "== pytest -m gpu" ; timeout 1500 python -m pytest tests -q -m gpu -x --tb=short > gpurun_out/pytest_gpu.log 2>&1 ; echo "pytest rc=$?" ; tail -30 gpurun_out/pytest_gpu.log
