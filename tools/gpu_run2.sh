set -x
cd $GRAFT_REPO_ROOT
python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "low_latency" 2>&1 | tail -25 > gpurun_out/r02_pytest_ll.log
python -m pytest tests/test_gpu_guards.py -x -q -m gpu -k "low_latency" 2>&1 | tail -25 >> gpurun_out/r02_pytest_ll.log
timeout 600 python tools/time_gru_ll.py > gpurun_out/r02_time_gru_ll.log 2>&1
tail -30 gpurun_out/r02_pytest_ll.log; cat gpurun_out/r02_time_gru_ll.log
