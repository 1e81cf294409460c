#!/bin/bash
# round 2: top-k with warp-sorted leader runs + counting order of the candidates: tests, kernel time under ncu
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r2_pytest28.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest28.log
tail -3 gpurun_out/r2_pytest28.log
timeout 300 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,dram__bytes_read.sum --clock-control none -k regex:topk_kernel --csv --log-file gpurun_out/r2_topk_ncu.csv python profiles/r2_topk_probe.py > gpurun_out/r2_topk_ncu.log 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(l for l in open('gpurun_out/r2_topk_ncu.csv') if not l.startswith('=='))]
h=rows[0]; mi=h.index('Metric Name'); vi=h.index('Metric Value'); ii=h.index('ID')
by={}
for r in rows[1:]:
    by.setdefault(r[ii],{})[r[mi]]=float(r[vi].replace(',',''))
vals=list(by.values())
for name,sl in (('k=20',vals[:19]),('k=100',vals[19:])):
    if sl:
        d=[v['gpu__time_duration.sum'] for v in sl]
        print(name, 'launches', len(sl), 'duration min %.1f us median %.1f us'%(min(d)/1e3, sorted(d)[len(d)//2]/1e3), 'warp instr %.1f M'%(sl[-1]['smsp__inst_executed.sum']/1e6), 'dram read %.0f MB'%(sl[-1]['dram__bytes_read.sum']/1e6))
PY
