#!/bin/bash
# round 2, GPU call 2: parity tests after the fix-up / prepare restructure, accuracy probe, bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2_pytest2.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest2.log
timeout 600 python tests/tools/accuracy_probe.py > gpurun_out/accuracy_r2.txt 2>&1
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_b.json 2> gpurun_out/r2_bench_b.err
tail -15 gpurun_out/r2_pytest2.log; cat gpurun_out/r2_bench_b.json | head -c 600
