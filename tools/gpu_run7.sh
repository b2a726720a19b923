cd $GRAFT_REPO_ROOT
python -m pytest tests -x -q -m gpu -k "not p5_full" 2>&1 | tail -8 > gpurun_out/r02_pytest_d.log
cat gpurun_out/r02_pytest_d.log
