#!/bin/bash
# round 2, final 8-GPU call: sharded tests, weak-scaling bench at N = 1, 2, 4, 8 (peer exchange) on one box, marks at N = 8
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_sharded.py -m gpu -q > gpurun_out/r2_pytest22.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest22.log
tail -3 gpurun_out/r2_pytest22.log
timeout 300 python bench.py --steps 20 --warmup 5 --no-extras --no-cpu-baseline > gpurun_out/r2_scale_n1.json 2> gpurun_out/r2_scale_n1.err
for n in 2 4 8; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 20 --warmup 5 > gpurun_out/r2_scale_n$n.json 2> gpurun_out/r2_scale_n$n.err
done
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29530 profiles/r2_step_marks.py > gpurun_out/r2_step_marks_n8.txt 2>&1; grep -v "^\*\|OMP_NUM\|^$" gpurun_out/r2_step_marks_n8.txt | tail -32
for f in r2_scale_n1 r2_scale_n2 r2_scale_n4 r2_scale_n8; do python - "$f" <<'PY'
import json,sys
f=sys.argv[1]
try:
    txt=open('gpurun_out/%s.json'%f).read()
    d=json.loads([l for l in txt.splitlines() if l.startswith('{')][-1])
    print(f, 'ms_per_step %.4f'%d['ms_per_step'], 'value %.4g'%d['value'], 'e2e %.4g'%d['e2e']['value'], 'launches', d['gpu_launches'])
    print('   parity', json.dumps(d.get('result',{}).get('parity'))[:200])
except Exception as e:
    print(f, 'FAILED', e); print(open('gpurun_out/%s.err'%f).read()[-2000:])
PY
done
