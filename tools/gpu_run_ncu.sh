cd $GRAFT_REPO_ROOT
CMD="python bench.py --steps 3 --warmup 3 --lean --no-cpu-baseline"
$CMD > gpurun_out/plain_bench.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches.csv $CMD > gpurun_out/ncu_launch.log 2>&1
tail -2 gpurun_out/ncu_launch.log
$CMD > gpurun_out/plain_bench2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"gru_bwd_tc_kernel|gru_fwd_tc_kernel|proj_fwd_tc|gru_dwhh_tc_kernel|proj_wgrad_tc_kernel|gd_prox_gc_kernel" -s 12 -c 12 -o gpurun_out/r02_prof_main $CMD > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log
