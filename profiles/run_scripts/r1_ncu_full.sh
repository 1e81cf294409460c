cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 500 ncu --set full --clock-control none --import-source on -k 'regex:distmat|group_|list_cap|pack_rows|rank_|topk' --launch-skip 16 --launch-count 16 -f -o gpurun_out/r1_full python profiles/ncu_targets.py > gpurun_out/ncu_run.log 2>&1
tail -3 gpurun_out/ncu_run.log | cut -c1-200
ls -la gpurun_out/r1_full.ncu-rep
