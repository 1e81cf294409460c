#!/bin/bash
# round 2: count kernel with software-pipelined loads: tests, kernel time under ncu, bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r2_pytest29.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest29.log
tail -3 gpurun_out/r2_pytest29.log
timeout 300 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,dram__bytes_read.sum,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:rank_count_warp --csv --log-file gpurun_out/r2_count_ncu.csv python profiles/r2_count_team_probe.py > gpurun_out/r2_count_ncu.log 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(l for l in open('gpurun_out/r2_count_ncu.csv') if not l.startswith('=='))]
h=rows[0]; mi=h.index('Metric Name'); vi=h.index('Metric Value'); ii=h.index('ID')
by={}
for r in rows[1:]:
    by.setdefault(r[ii],{})[r[mi]]=float(r[vi].replace(',',''))
vals=list(by.values())
per=len(vals)//5
for t,sl in zip((1,2,4,8,1),[vals[i*per:(i+1)*per] for i in range(5)]):
    d=[v['gpu__time_duration.sum'] for v in sl]
    print('team',t,'launches',len(sl),'duration min %.1f us median %.1f us'%(min(d)/1e3, sorted(d)[len(d)//2]/1e3),'warp instr %.1f M'%(sl[-1]['smsp__inst_executed.sum']/1e6),'issue active %.0f %%'%sl[-1]['smsp__issue_active.avg.pct_of_peak_sustained_active'])
PY
timeout 300 python bench.py --no-extras --no-cpu-baseline --no-parity > gpurun_out/r2_b29.json 2> gpurun_out/r2_b29.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r2_b29.json') if l.startswith('{')][-1])
print('bench ms_per_step %.4f'%d['ms_per_step'], json.dumps(d['stage_ms']))
PY
