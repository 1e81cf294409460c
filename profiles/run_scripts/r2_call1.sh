#!/bin/bash
# round 2, GPU call 1: parity tests, accuracy probe, contraction sweep, first bench line
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest1.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest1.log
timeout 600 python tests/tools/accuracy_probe.py > gpurun_out/accuracy_r2.txt 2>&1
timeout 600 python profiles/r2_gemm_sweep.py > gpurun_out/r2_gemm_sweep.txt 2>&1
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench_a.json 2> gpurun_out/r2_bench_a.err
tail -5 gpurun_out/r2_pytest1.log; tail -3 gpurun_out/r2_gemm_sweep.txt; cat gpurun_out/r2_bench_a.json | head -c 1500
