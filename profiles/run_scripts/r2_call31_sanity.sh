#!/bin/bash
# round 2, last call: the driver's round-end sequence on the final tree (tests, smoke, bench)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r2_pytest31.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest31.log
tail -3 gpurun_out/r2_pytest31.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 300 python bench.py --gpus 1 --steps 20 --warmup 5 --no-extras > gpurun_out/r2_b31.json 2> gpurun_out/r2_b31.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r2_b31.json') if l.startswith('{')][-1])
print('bench ms_per_step %.4f'%d['ms_per_step'], 'e2e ms %.3f'%d['e2e']['ms_per_step'], 'launches', d['gpu_launches'], json.dumps(d['stage_ms']), d['roofline_rank_count']['frac'], json.dumps(d['result']['parity'])[:200])
PY
