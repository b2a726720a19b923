set -x
cd $GRAFT_REPO_ROOT
nvidia-smi -L
python -m pytest tests -x -q -m gpu -k "not p5_full and not guard" 2>&1 | tail -25 > gpurun_out/r02_pytest_a.log
python -m pytest tests/test_gpu_guards.py -q -m gpu 2>&1 | tail -40 > gpurun_out/r02_pytest_guards.log
python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_a.json 2> gpurun_out/r02_bench_a.err
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r02_bench_ref_a.json 2> gpurun_out/r02_bench_ref_a.err
tail -3 gpurun_out/r02_pytest_a.log gpurun_out/r02_pytest_guards.log; tail -c 600 gpurun_out/r02_bench_a.err
