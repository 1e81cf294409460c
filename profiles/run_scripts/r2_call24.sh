#!/bin/bash
# round 2: the bench line with the final timing harness (defaults = the driver's 20 / 5)
mkdir -p gpurun_out
( time timeout 900 python bench.py > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err ) 2>&1 | grep real
timeout 300 python bench.py --no-extras --no-cpu-baseline --no-parity > gpurun_out/r2_b24_a.json 2> gpurun_out/r2_b24_a.err
timeout 300 python bench.py --no-extras --no-cpu-baseline --no-parity > gpurun_out/r2_b24_b.json 2> gpurun_out/r2_b24_b.err
for f in r2_bench_n1 r2_b24_a r2_b24_b; do python - "$f" <<'PY'
import json,sys
f=sys.argv[1]
try:
    txt=open('gpurun_out/%s.json'%f).read()
    d=json.loads([l for l in txt.splitlines() if l.startswith('{')][-1])
    print(f, 'ms_per_step %.4f'%d['ms_per_step'], 'value %.4g'%d['value'], 'e2e %.4g'%d['e2e']['value'], 'launches', d['gpu_launches'], d['clocks'])
except Exception as e:
    print(f, 'FAILED', e); print(open('gpurun_out/%s.err'%f).read()[-2000:])
PY
done
