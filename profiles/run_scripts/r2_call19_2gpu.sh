#!/bin/bash
# round 2: queries packed in the gallery call, grouping joined before the gather, warp-per-query metrics: tests, bench and marks at N = 1, 2
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/r2_pytest19.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest19.log
tail -12 gpurun_out/r2_pytest19.log
timeout 300 python bench.py --steps 20 --warmup 5 --no-extras --no-cpu-baseline > gpurun_out/r2_b19_n1.json 2> gpurun_out/r2_b19_n1.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2_b19_n2.json 2> gpurun_out/r2_b19_n2.err
for f in r2_b19_n1 r2_b19_n2; do python - "$f" <<'PY'
import json,sys
f=sys.argv[1]
try:
    txt=open('gpurun_out/%s.json'%f).read()
    d=json.loads([l for l in txt.splitlines() if l.startswith('{')][-1])
    print(f, 'ms_per_step %.4f'%d['ms_per_step'], 'value %.4g'%d['value'], 'e2e ms %.4g'%d['e2e']['ms_per_step'], 'launches', d['gpu_launches'], d['clocks'])
    print('   parity', json.dumps(d.get('result',{}).get('parity'))[:400])
except Exception as e:
    print(f, 'FAILED', e); print(open('gpurun_out/%s.err'%f).read()[-2000:])
PY
done
timeout 200 python profiles/r2_step_marks.py > gpurun_out/r2_step_marks_n1.txt 2>&1; cat gpurun_out/r2_step_marks_n1.txt | tail -25
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29532 profiles/r2_step_marks.py > gpurun_out/r2_step_marks_n2.txt 2>&1; grep -v "^\*\|OMP_NUM\|^$" gpurun_out/r2_step_marks_n2.txt | tail -40
timeout 200 python profiles/r2_sharded_step_probe.py 2>&1 | grep "^rank"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 profiles/r2_sharded_step_probe.py 2>&1 | grep "^rank"
