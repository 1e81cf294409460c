#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r2_pytest15.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest15.log
tail -5 gpurun_out/r2_pytest15.log
timeout 300 python bench.py --steps 20 --warmup 5 --no-extras --no-cpu-baseline > gpurun_out/r2_b15_n1.json 2> gpurun_out/r2_b15_n1.err
timeout 300 python bench.py --no-extras --no-cpu-baseline > gpurun_out/r2_b15_n1_default.json 2> gpurun_out/r2_b15_n1_default.err
timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:distmat|group_|list_cap|pack_rows|center_|rank_|topk' --launch-skip 21 --launch-count 21 -f -o gpurun_out/r2_full python profiles/ncu_targets.py > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
for f in r2_b15_n1 r2_b15_n1_default; do python - "$f" <<'PY'
import json,sys
f=sys.argv[1]
try:
    txt=open('gpurun_out/%s.json'%f).read()
    d=json.loads([l for l in txt.splitlines() if l.startswith('{')][-1])
    print(f, 'ms_per_step %.4f'%d['ms_per_step'], 'value %.4g'%d['value'], 'e2e %.4g'%d['e2e']['value'], 'launches', d['gpu_launches'], json.dumps(d['stage_ms']), d['clocks'])
except Exception as e:
    print(f, 'FAILED', e); print(open('gpurun_out/%s.err'%f).read()[-2000:])
PY
done
