set -x
cd $GRAFT_REPO_ROOT
python tools/prof_ll.py > gpurun_out/plain_ll.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:gru_.*_ll_kernel -s 2 -c 2 -o gpurun_out/r02_prof_ll python tools/prof_ll.py > gpurun_out/ncu_ll.log 2>&1
tail -5 gpurun_out/ncu_ll.log
