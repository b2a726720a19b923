cd $GRAFT_REPO_ROOT
python -m pytest tests/test_gpu_train.py -x -q -m gpu -k "flow_iteration or cuda_graph_replay or host_fed or p3_short or p4_train" 2>&1 | tail -8 > gpurun_out/r02_pytest_flow.log
cat gpurun_out/r02_pytest_flow.log
for G in auto "74,26" "50,50" "37,37,26" ; do
  CRVAE_GROUPS=$G python bench.py --steps 200 --warmup 10 --lean --no-cpu-baseline > gpurun_out/r02_bench_flow_$G.json 2> gpurun_out/r02_bench_flow_$G.err
  python -c "
import json,sys
d=json.loads(open('gpurun_out/r02_bench_flow_$G.json').read().strip().splitlines()[-1])
print('$G', d['ms_per_step'], d['e2e']['ms_per_step'], d['flow'], d['loss_after_timed'])"
done
CRVAE_FLOW=0 python bench.py --steps 200 --warmup 10 --lean --no-cpu-baseline > gpurun_out/r02_bench_noflow.json 2>/dev/null
python -c "
import json
d=json.loads(open('gpurun_out/r02_bench_noflow.json').read().strip().splitlines()[-1])
print('noflow', d['ms_per_step'], d['e2e']['ms_per_step'], d['flow'], d['loss_after_timed'])"
