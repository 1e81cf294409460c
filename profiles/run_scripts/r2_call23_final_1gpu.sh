#!/bin/bash
# round 2, final 1-GPU call: all tests, the full bench line (parity + extras + cpu baseline), reference arm, marks,
# ncu launch list + full capture of every kernel of the step
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r2_pytest23.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest23.log
tail -4 gpurun_out/r2_pytest23.log
( time timeout 900 python bench.py > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err ) 2>&1 | grep real
( time timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_reference_arm.json 2> gpurun_out/r2_bench_ref.err ) 2>&1 | grep real
timeout 200 python profiles/r2_step_marks.py > gpurun_out/r2_step_marks_n1.txt 2>&1; tail -14 gpurun_out/r2_step_marks_n1.txt
timeout 200 python profiles/r2_sharded_step_probe.py 2>&1 | grep "^rank" | tee gpurun_out/r2_step_probe_n1.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline --no-parity > gpurun_out/ncu_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:distmat|group_|list_cap|pack_rows|center_|rank_|topk' --launch-skip 21 --launch-count 21 -f -o gpurun_out/r2_full python profiles/ncu_targets.py > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log; ls -la gpurun_out/r2_full.ncu-rep
python - <<'PY'
import json
txt=open('gpurun_out/r2_bench_n1.json').read()
d=json.loads([l for l in txt.splitlines() if l.startswith('{')][-1])
print('ms_per_step', d['ms_per_step'], 'value', d['value'], 'e2e', d['e2e'], 'clocks', d['clocks'])
print('roofline', json.dumps(d['roofline'])[:500])
print('count', json.dumps(d['roofline_rank_count'])[:400])
print('stage', d['stage_ms'])
print('cpu', d['cpu_baseline'])
print('parity', json.dumps(d['result']['parity'])[:500])
for k,v in (d.get('extras') or {}).items(): print('  ', k, json.dumps(v)[:900])
print(open('gpurun_out/r2_bench_reference_arm.json').read()[:900])
PY
