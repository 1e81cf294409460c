#!/bin/bash
mkdir -p gpurun_out
timeout 200 python profiles/r2_sharded_step_probe.py > gpurun_out/r2_step_probe_n1.txt 2>&1; tail -2 gpurun_out/r2_step_probe_n1.txt
for n in 2 8; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2953$n profiles/r2_sharded_step_probe.py > gpurun_out/r2_step_probe_n$n.txt 2>&1; grep "^rank" gpurun_out/r2_step_probe_n$n.txt
done
IEEE_B200_TRACE=1 IEEE_B200_ONE_CALL=0 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 profiles/sharded_timeline.py > gpurun_out/r2_sharded_timeline_n8.txt 2>&1; tail -40 gpurun_out/r2_sharded_timeline_n8.txt
