cd $GRAFT_REPO_ROOT
python bench.py --steps 1000 --warmup 10 > gpurun_out/r02c_bench_main.json 2> gpurun_out/r02c_bench_main.err
python bench.py --p 10 --steps 200 --warmup 10 --lean > gpurun_out/r02c_bench_p10.json 2> gpurun_out/r02c_bench_p10.err
python bench.py --p 1000 --T 2000 --steps 20 --warmup 3 --lean --no-cpu-baseline > gpurun_out/r02c_bench_p1000_n1.json 2> gpurun_out/r02c_bench_p1000_n1.err
python tools/bench_config4.py > gpurun_out/r02c_bench_config4.json 2> gpurun_out/r02c_bench_config4.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02c_bench_reference_arm.json 2> gpurun_out/r02c_bench_reference_arm.err
for f in main p10 p1000_n1; do python -c "
import json
d=json.loads(open('gpurun_out/r02c_bench_$f.json').read().strip().splitlines()[-1])
print('$f', d['ms_per_step'], d['value'], d['e2e']['ms_per_step'], d['flow'], d['roofline']['kernel'], round(d['roofline']['frac'],3), d.get('phase2',{}).get('ms_per_step'))"; done
cat gpurun_out/r02c_bench_config4.json
tail -c 600 gpurun_out/r02c_bench_reference_arm.json
