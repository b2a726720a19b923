"""Phase-2 iteration time at p = 100 (masked-dense), several repetitions: CUDA events around graph replays."""
import os, sys, json
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vae_connexe_b200 as V
import bench
Xb = bench.make_batch(100, 1000, 256).cuda()
for rep in range(3):
    r = bench.time_phase2(V, Xb, 100, 256, 200, packed=False)
    print(json.dumps({"CRVAE_MMA": os.environ.get("CRVAE_MMA", "small"), "rep": rep, "ms_per_step": r["ms_per_step"]}), flush=True)
