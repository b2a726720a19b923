set -x
cd $GRAFT_REPO_ROOT
python -m pytest tests -x -q -m gpu -k "not p5_full" 2>&1 | tail -15 > gpurun_out/r02_pytest_b.log
python bench.py --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/r02_bench_b.json 2> gpurun_out/r02_bench_b.err
tail -4 gpurun_out/r02_pytest_b.log; tail -c 300 gpurun_out/r02_bench_b.err
