#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_engine.py -m gpu -q > gpurun_out/r2_pytest9.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest9.log
tail -5 gpurun_out/r2_pytest9.log
timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:chunked_kernel|fused_prepass|fused_recount' --launch-skip 9 --launch-count 3 -f -o gpurun_out/r2_fused_full python profiles/r2_fused_probe.py > gpurun_out/ncu_fused_full.log 2>&1
tail -3 gpurun_out/ncu_fused_full.log; ls -la gpurun_out/r2_fused_full.ncu-rep
