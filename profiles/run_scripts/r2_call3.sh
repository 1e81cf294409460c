#!/bin/bash
# round 2, GPU call 3: isolate the epilogue cost of the near-duplicate detection; tests; bench; step timeline
mkdir -p gpurun_out
timeout 600 python profiles/r2_gemm_sweep.py --quick > gpurun_out/r2_gemm_sweep_b.txt 2>&1
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r2_pytest3.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest3.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_c.json 2> gpurun_out/r2_bench_c.err
timeout 300 python profiles/r2_step_timeline.py > gpurun_out/r2_step_timeline.txt 2>&1
cat gpurun_out/r2_gemm_sweep_b.txt; tail -4 gpurun_out/r2_pytest3.log; cat gpurun_out/r2_bench_c.json | head -c 400; cat gpurun_out/r2_step_timeline.txt
