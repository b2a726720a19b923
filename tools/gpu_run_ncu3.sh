cd $GRAFT_REPO_ROOT
export P=13 T=10 B=256
python tools/prof_mma.py > gpurun_out/prof_mma_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"gru_fwd_mma|gru_bwd_mma" -s 2 -c 2 -o gpurun_out/r02c_prof_mma_p13 -f python tools/prof_mma.py > gpurun_out/prof_mma_ncu.log 2>&1
tail -1 gpurun_out/prof_mma_ncu.log
export P=100
python tools/prof_mma.py > gpurun_out/prof_mma_plain.log 2>&1 && ncu --set full --clock-control none -k regex:"gru_fwd_mma|gru_bwd_mma" -s 2 -c 2 -o gpurun_out/r02c_prof_mma_p100 -f python tools/prof_mma.py > gpurun_out/prof_mma_ncu.log 2>&1
tail -1 gpurun_out/prof_mma_ncu.log
