cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_c1_real_model.py tests/test_visrank.py tests/test_gpu_distance.py tests/test_gpu_engine.py -m gpu -x -q 2>&1 | tail -3
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 profiles/sharded_timeline.py 2>&1 | grep -E "trace|world"
