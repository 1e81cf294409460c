#!/bin/bash
# round 2, GPU call 6: the fused-count path -- parity tests first (bounded), then the bench with it on and off
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_engine.py -m gpu -q -x > gpurun_out/r2_pytest6a.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest6a.log
tail -30 gpurun_out/r2_pytest6a.log
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2_pytest6.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest6.log
tail -6 gpurun_out/r2_pytest6.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-extras --no-cpu-baseline > gpurun_out/r2_bench_fused.json 2> gpurun_out/r2_bench_fused.err
IEEE_B200_FUSED=0 timeout 600 python bench.py --steps 20 --warmup 5 --no-extras --no-cpu-baseline > gpurun_out/r2_bench_staged.json 2> gpurun_out/r2_bench_staged.err
for f in r2_bench_fused r2_bench_staged; do python - "$f" <<'PY'
import json,sys
f=sys.argv[1]
try:
    txt=open('gpurun_out/%s.json'%f).read()
    d=json.loads([l for l in txt.splitlines() if l.startswith('{')][-1])
    print(f, 'ms_per_step', d['ms_per_step'], 'value', d['value'], 'launches', d['gpu_launches'], d.get('count_path'))
    print('   parity', json.dumps(d.get('result',{}).get('parity'))[:600])
except Exception as e:
    print(f, 'FAILED', e); print(open('gpurun_out/%s.err'%f).read()[-2500:])
PY
done
