#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_engine.py tests/test_gpu_sharded.py -m gpu -q > gpurun_out/r2_pytest7.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest7.log
tail -12 gpurun_out/r2_pytest7.log
timeout 300 python profiles/r2_fused_probe.py --events > gpurun_out/r2_fused_probe.txt 2>&1; cat gpurun_out/r2_fused_probe.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k 'regex:fused|distmat|pack_rows|center|group|rank_' --launch-skip 60 --launch-count 40 --csv --log-file gpurun_out/r2_fused_launches.csv python profiles/r2_fused_probe.py > gpurun_out/ncu_fused.log 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/r2_fused_launches.csv')) if len(r)>10]
hdr=rows[0]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value')
for r in rows[1:41]: print(r[ki][:70], r[vi])
PY
