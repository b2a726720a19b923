cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -4 > gpurun_out/r02c_pytest_gpu.log
cat gpurun_out/r02c_pytest_gpu.log
bash tools/gpu_run13.sh
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
