set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" 
tail -5 gpurun_out/pytest_gpu.log
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_onecall.log 2>&1; tail -1 gpurun_out/bench_onecall.log | cut -c1-600
IEEE_B200_ONE_CALL=0 timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_staged.log 2>&1; tail -1 gpurun_out/bench_staged.log | cut -c1-300
timeout 300 python profiles/prof_kernels.py --reps 10 2>&1 | grep -E "rank_|group|pack" 
timeout 200 python profiles/count_large.py 8192
IEEE_B200_COUNT_WARP_MAX_G=200000 timeout 200 python profiles/count_large.py 8192
