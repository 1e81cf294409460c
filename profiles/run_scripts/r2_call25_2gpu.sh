#!/bin/bash
# round 2: the final timing harness under torchrun (N = 2)
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2_b25_n2.json 2> gpurun_out/r2_b25_n2.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/r2_b25_ref_n2.json 2> gpurun_out/r2_b25_ref_n2.err
python - <<'PY'
import json
for f in ('r2_b25_n2','r2_b25_ref_n2'):
    try:
        txt=open('gpurun_out/%s.json'%f).read()
        d=json.loads([l for l in txt.splitlines() if l.startswith('{')][-1])
        print(f, 'ms_per_step %.4f'%d['ms_per_step'], 'value %.4g'%d['value'], 'e2e %.4g'%d['e2e']['value'], d.get('gpu_launches'), d.get('clocks'), json.dumps(d.get('result',{}).get('parity'))[:200])
    except Exception as e:
        print(f, 'FAILED', e); print(open('gpurun_out/%s.err'%f).read()[-2000:])
PY
