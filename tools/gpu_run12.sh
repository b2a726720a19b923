cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -15 > gpurun_out/r02_pytest_gpu_mma.log
cat gpurun_out/r02_pytest_gpu_mma.log
python bench.py --steps 200 --warmup 10 > gpurun_out/r02_bench_main_mma.json 2> gpurun_out/r02_bench_main_mma.err
python -c "
import json
d=json.loads(open('gpurun_out/r02_bench_main_mma.json').read().strip().splitlines()[-1])
print('main', d['ms_per_step'], d['value'], d['e2e']['ms_per_step'], d['flow'], d['roofline']['kernel'], round(d['roofline']['frac'],3))
for k,v in sorted(d['stages_ms'].items(), key=lambda kv:-kv[1]): print('  %-28s %.4f' % (k,v))
"
