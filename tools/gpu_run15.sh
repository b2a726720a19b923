cd $GRAFT_REPO_ROOT
bash tools/gpu_run13.sh
CMD="python bench.py --steps 3 --warmup 3 --lean --no-cpu-baseline"
$CMD > gpurun_out/plain_bench.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02c_launches.csv $CMD > gpurun_out/ncu_launch.log 2>&1
tail -c 300 gpurun_out/ncu_launch.log
