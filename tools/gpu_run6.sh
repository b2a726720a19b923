cd $GRAFT_REPO_ROOT
python -m pytest tests/test_gpu_train.py -x -q -m gpu -k "p100_matches or packed or check_block_generation or generation_matches" 2>&1 | tail -5 > gpurun_out/r02_pytest_c.log
python tools/time_phase2.py > gpurun_out/r02_time_phase2.log 2>&1
cat gpurun_out/r02_pytest_c.log gpurun_out/r02_time_phase2.log
