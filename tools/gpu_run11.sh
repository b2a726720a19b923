cd $GRAFT_REPO_ROOT
python -m pytest tests/test_gpu_train.py -x -q -m gpu -k "mixture or teacher_forcing or config4 or generic_vrae" 2>&1 | tail -5 > gpurun_out/r02_pytest_e.log
cat gpurun_out/r02_pytest_e.log
python tools/bench_config4.py 2>/dev/null | tail -1 > gpurun_out/r02_bench_config4.json; cat gpurun_out/r02_bench_config4.json
