"""torchrun worker: row-sharded CUDA evaluation over NCCL vs the same evaluation on one GPU."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gpz_b200 import _lib as L  # noqa: E402
from gpz_b200 import synth  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
ok = True
for method, n, d, m in (("VC", 20011, 5, 150), ("VD", 9000, 3, 40)):
    X, Y = synth.make_data(n, d, seed=1)
    theta = synth.perturb_theta(synth.make_theta0(X, Y, method, m, het=True, seed=2), 0.05, 3)
    tr = np.arange(n) % 5 != 0
    lo, hi = (n * rank) // world, (n * (rank + 1)) // world
    gm = L.make_model(d, 1, m, method, True)
    ctx = L.Context(gm, X[lo:hi], Y[lo:hi], None, None, tr[lo:hi], ~tr[lo:hi], device=local)
    uid = [L.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    ctx.comm_init(rank, world, uid[0])
    f, g, st = ctx.eval(theta)
    nl, w, iS = ctx.fit(theta)
    ctx.close()
    if rank == 0:
        one = L.Context(gm, X, Y, None, None, tr, ~tr, device=local)
        f1, g1, st1 = one.eval(theta)
        nl1, w1, iS1 = one.fit(theta)
        one.close()
        e = [abs(f - f1) / abs(f1), np.max(np.abs(g - g1)) / np.max(np.abs(g1)), np.max(np.abs(w - w1)) / np.max(np.abs(w1)),
             abs(nl[0, 0] - nl1[0, 0]) / abs(nl1[0, 0])] + [abs(st[k] - st1[k]) for k in st]
        print(method, "sharded vs single:", e, flush=True)
        ok = ok and max(e) < 1e-10
    dist.barrier()
if rank == 0:
    print("MGPU_OK" if ok else "MGPU_FAIL", flush=True)
dist.destroy_process_group()
sys.exit(0 if ok else 1)
