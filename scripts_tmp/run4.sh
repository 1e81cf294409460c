cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_rank.py tests/test_gpu_sharded.py tests/test_gpu_engine.py -m gpu -x -q 2>&1 | tail -3
for n in 1 2 4 8; do timeout 200 python profiles/count_sharded.py $n; done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:rank_count -c 1 -o gpurun_out/count_n1b python profiles/count_sharded.py 1 --reps 1 --width 0 > gpurun_out/ncu_n1b.log 2>&1
