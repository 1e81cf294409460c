#!/bin/bash
# round 2, 1 GPU: tests with the team-size count kernel and the staged gather; count team probe; bench + marks
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/r2_pytest21.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest21.log
tail -4 gpurun_out/r2_pytest21.log
timeout 200 python profiles/r2_count_team_probe.py > gpurun_out/r2_count_team_probe.txt 2>&1; tail -6 gpurun_out/r2_count_team_probe.txt
for t in 1 2 4; do
IEEE_B200_COUNT_WPQ=$t timeout 300 python bench.py --steps 20 --warmup 5 --no-extras --no-cpu-baseline --no-parity > gpurun_out/r2_b21_team$t.json 2> gpurun_out/r2_b21_team$t.err
python - "r2_b21_team$t" <<'PY'
import json,sys
f=sys.argv[1]
try:
    txt=open('gpurun_out/%s.json'%f).read()
    d=json.loads([l for l in txt.splitlines() if l.startswith('{')][-1])
    print(f, 'ms_per_step %.4f'%d['ms_per_step'], 'value %.4g'%d['value'], json.dumps(d['stage_ms']))
except Exception as e:
    print(f, 'FAILED', e); print(open('gpurun_out/%s.err'%f).read()[-2000:])
PY
done
