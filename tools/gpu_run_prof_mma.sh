cd $GRAFT_REPO_ROOT
export P=${P:-100} T=${T:-10} B=${B:-256}
python tools/prof_mma.py > gpurun_out/prof_mma_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"gru_fwd_mma|gru_bwd_mma" -s 2 -c 2 -o gpurun_out/prof_mma python tools/prof_mma.py > gpurun_out/prof_mma_ncu.log 2>&1
tail -3 gpurun_out/prof_mma_ncu.log
