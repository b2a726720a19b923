cd $GRAFT_REPO_ROOT
python -m pytest tests/test_gpu_train.py -x -q -m gpu -k "full_config4_matches_oracle" 2>&1 | grep -E "assert|Error|rel|^E " | head -20 > gpurun_out/r02_cfg4.log
CRVAE_LL=0 python -m pytest tests/test_gpu_train.py -x -q -m gpu -k "full_config4_matches_oracle" 2>&1 | tail -3 >> gpurun_out/r02_cfg4.log
cat gpurun_out/r02_cfg4.log
