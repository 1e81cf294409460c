#!/bin/bash
# round 2, GPU call 4 (2 GPUs): all GPU tests incl. the peer exchange, then the weak-scaling bench at N=1 and N=2
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r2_pytest4.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest4.log
tail -25 gpurun_out/r2_pytest4.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2_bench_n1_d.json 2> gpurun_out/r2_bench_n1_d.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2_bench_n2_d.json 2> gpurun_out/r2_bench_n2_d.err
IEEE_B200_EXCHANGE=nccl timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2_bench_n2_nccl.json 2> gpurun_out/r2_bench_n2_nccl.err
for f in r2_bench_n1_d r2_bench_n2_d r2_bench_n2_nccl; do python - "$f" <<'PY'
import json,sys
f=sys.argv[1]
try:
    d=json.load(open('gpurun_out/%s.json'%f)); print(f, 'ms_per_step', d['ms_per_step'], 'value', d['value'], 'e2e', d['e2e']['value'], 'mAP', d['result']['mAP'])
except Exception as e:
    print(f, 'FAILED', e); print(open('gpurun_out/%s.err'%f).read()[-1500:])
PY
done
