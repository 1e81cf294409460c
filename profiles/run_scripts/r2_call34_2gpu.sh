#!/bin/bash
# round 2, last multi-GPU check of the final tree: the 2-GPU tests and the N = 2 bench line
mkdir -p gpurun_out
timeout 100 python -m pytest tests/test_gpu_sharded.py -m gpu -q > gpurun_out/r2_pytest34.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest34.log
tail -3 gpurun_out/r2_pytest34.log
timeout 100 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2_b34_n2.json 2> gpurun_out/r2_b34_n2.err
python - <<'PY'
import json
try:
    d=json.loads([l for l in open('gpurun_out/r2_b34_n2.json') if l.startswith('{')][-1])
    print('N=2 ms_per_step %.4f'%d['ms_per_step'], 'value %.4g'%d['value'], 'e2e %.4g'%d['e2e']['value'], json.dumps(d['result']['parity'])[:160])
except Exception as e:
    print('FAILED', e); print(open('gpurun_out/r2_b34_n2.err').read()[-1500:])
PY
