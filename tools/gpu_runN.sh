cd $GRAFT_REPO_ROOT
N=$1
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 200 --warmup 10 --lean > gpurun_out/r02c_bench_n$N.json 2> gpurun_out/r02c_bench_n$N.err
tail -c 400 gpurun_out/r02c_bench_n$N.err
python -c "
import json
d=json.loads(open('gpurun_out/r02c_bench_n$N.json').read().strip().splitlines()[-1])
print('N=$N', d['ms_per_step'], d['e2e']['ms_per_step'], d['flow'], d['loss_after_timed'], d['parity_vs_n1'], d['l2_flush_between_steps'])
for k,v in sorted(d['stages_ms'].items(), key=lambda kv:-kv[1]): print('  %-28s %.4f' % (k,v))
"
