#!/bin/bash
# round 2, GPU call 5 (2 GPUs): tests, the full N=1 bench line (parity + extras), N=2 with the one-call peer path, reference arm
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/r2_pytest5.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest5.log
tail -5 gpurun_out/r2_pytest5.log
( time timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_n1_e.json 2> gpurun_out/r2_bench_n1_e.err ) 2>&1 | grep real
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2_bench_n2_e.json 2> gpurun_out/r2_bench_n2_e.err
( time timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_ref_e.json 2> gpurun_out/r2_bench_ref_e.err ) 2>&1 | grep real
for f in r2_bench_n1_e r2_bench_n2_e r2_bench_ref_e; do python - "$f" <<'PY'
import json,sys
f=sys.argv[1]
try:
    txt=open('gpurun_out/%s.json'%f).read()
    d=json.loads([l for l in txt.splitlines() if l.startswith('{')][-1])
    print(f, 'ms_per_step', d['ms_per_step'], 'value', d['value'], 'e2e', d['e2e']['value'])
    print('   parity', json.dumps(d.get('result',{}).get('parity'))[:900])
    for k,v in (d.get('extras') or {}).items(): print('   ', k, json.dumps(v)[:1200])
except Exception as e:
    print(f, 'FAILED', e); print(open('gpurun_out/%s.err'%f).read()[-2500:])
PY
done
