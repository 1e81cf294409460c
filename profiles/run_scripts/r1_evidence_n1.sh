cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/gpu.txt
timeout 400 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_r1_n1.log 2>gpurun_out/bench_r1_n1.err; tail -1 gpurun_out/bench_r1_n1.log | cut -c1-200
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_r1_ref.log 2>/dev/null; tail -1 gpurun_out/bench_r1_ref.log | cut -c1-200
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r1_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r1_launch_run.log 2>&1
timeout 500 ncu --set full --clock-control none --import-source on -k regex:ieee --launch-skip 16 --launch-count 16 -f -o gpurun_out/r1_full python profiles/ncu_targets.py > gpurun_out/ncu_run.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:rank_count --launch-skip 1 --launch-count 1 -f -o gpurun_out/r1_count_long python profiles/count_large.py 8192 > gpurun_out/ncu_count_long.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -3
