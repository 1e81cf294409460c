cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 200 python profiles/count_large.py 8192
timeout 200 python profiles/count_large.py 2048
for n in 1 8; do timeout 200 python profiles/count_sharded.py $n; done
