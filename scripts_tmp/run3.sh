cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for n in 1 2 4 8; do timeout 200 python profiles/count_sharded.py $n; done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:rank_count -c 2 -o gpurun_out/count_n2 python profiles/count_sharded.py 2 --reps 1 > gpurun_out/ncu_n2.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:rank_count -c 2 -o gpurun_out/count_n8 python profiles/count_sharded.py 8 --reps 1 > gpurun_out/ncu_n8.log 2>&1
ls -la gpurun_out/*.ncu-rep
