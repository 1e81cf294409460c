cd $GRAFT_REPO_ROOT
for n in 1 2 4 8; do IEEE_B200_COUNT_WARP_MAX_G=100 timeout 200 python profiles/count_sharded.py $n; done
