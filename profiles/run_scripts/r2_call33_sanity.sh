#!/bin/bash
# round 2, last call: all GPU tests + smoke on the final tree
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2_pytest33.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest33.log
tail -6 gpurun_out/r2_pytest33.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
