"""Phase-2 iteration time (CRVAE on the pruned Lorenz-96 ring + VRAE4E error compensation, CRVAE_lorenz96.py:609-643), B=256:
masked-dense vs gather-packed storage of the ragged heads, CUDA events around graph replays."""
import os, sys, json
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vae_connexe_b200 as V
import bench

for p, T in ((100, 1000), (1000, 2000)):
    Xb = bench.make_batch(p, T, 256).cuda()
    for packed in (False, True):
        if p == 1000 and not packed and os.environ.get("SKIP_DENSE_1000"):
            continue
        r = bench.time_phase2(V, Xb, p, 256, 50, packed=packed)
        print(json.dumps({"p": p, "packed": packed, **{k: r[k] for k in ("ms_per_step", "value", "w_ih_bytes", "loss", "loss_e")}}), flush=True)
        torch.cuda.empty_cache()
