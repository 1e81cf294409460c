cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513 profiles/sharded_timeline.py 2>&1 | grep -E "trace|world" 
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 8 --steps 20 --warmup 3 > gpurun_out/bench_n8.log 2>&1; tail -1 gpurun_out/bench_n8.log | cut -c1-260; grep -o '"e2e": {[^}]*}' gpurun_out/bench_n8.log | tail -1
