#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/parity_report.txt
echo "== conv_tc isolated" ; timeout 600 python -m pytest tests/test_gpu_conv_tc.py -q --tb=short > gpurun_out/pytest_tc.log 2>&1 ; echo "rc=$?" ; tail -40 gpurun_out/pytest_tc.log
echo "== full gpu suite" ; timeout 1500 python -m pytest tests -q -m gpu --tb=short --deselect tests/test_gpu_conv_tc.py > gpurun_out/pytest_gpu.log 2>&1 ; echo "rc=$?" ; tail -30 gpurun_out/pytest_gpu.log
