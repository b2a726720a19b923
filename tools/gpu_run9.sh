cd $GRAFT_REPO_ROOT
for G in "50,50" "60,40" "34,33,33" "25,25,25,25" "50,25,25" "20,20,20,20,20" "100"; do
  CRVAE_GROUPS=$G python bench.py --steps 200 --warmup 10 --lean --no-cpu-baseline > gpurun_out/r02_bench_flow_$G.json 2> gpurun_out/r02_bench_flow_$G.err
  python -c "
import json,sys
d=json.loads(open('gpurun_out/r02_bench_flow_$G.json').read().strip().splitlines()[-1])
print('$G', round(d['ms_per_step'],4), round(d['e2e']['ms_per_step'],4), d['flow'], d['loss_after_timed'])"
done
