cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 110 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 profiles/c4_sharded.py --precision f16x3 --check 256 > gpurun_out/c4_f16x3.log 2>&1; tail -1 gpurun_out/c4_f16x3.log | cut -c1-700
timeout 90 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29522 profiles/c4_sharded.py --precision bf16 --check 0 > gpurun_out/c4_bf16.log 2>&1; tail -1 gpurun_out/c4_bf16.log | cut -c1-700
