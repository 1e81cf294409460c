#!/bin/bash
# round 2: top-k select path with float leaders and vector loads: all tests, timing probe
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r2_pytest26.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest26.log
tail -4 gpurun_out/r2_pytest26.log
timeout 200 python profiles/r2_topk_probe.py > gpurun_out/r2_topk_probe.txt 2>&1; tail -3 gpurun_out/r2_topk_probe.txt
