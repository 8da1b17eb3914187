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
    # the device-resident optimiser on the sharded objective: every rank takes the same decisions (same f, g after the
    # all-reduce), so all ranks return the same theta bit for bit
    th_t, best_t, bv_t, info_t = ctx.train(theta, theta, -np.inf, max_iter=8, max_attempts=3.0, training_only=0)
    sig = torch.tensor([float(np.sum(th_t)), float(np.sum(best_t)), bv_t, float(info_t["fun_evals"])], dtype=torch.float64,
                       device=f"cuda:{local}")
    sigs = [torch.empty_like(sig) for _ in range(world)]
    dist.all_gather(sigs, sig)
    same = all(torch.equal(q, sigs[0]) for q in sigs)
    ctx.close()
    if rank == 0:
        one = L.Context(gm, X, Y, None, None, tr, ~tr, device=local)
        f1, g1, st1 = one.eval(theta)
        nl1, w1, iS1 = one.fit(theta)
        th_1, best_1, bv_1, info_1 = one.train(theta, theta, -np.inf, max_iter=8, max_attempts=3.0, training_only=0)
        one.close()
        et = [abs(info_t["f"] - info_1["f"]) / abs(info_1["f"]), np.max(np.abs(th_t - th_1)) / np.max(np.abs(th_1))]
        print(method, "sharded train vs single:", et, "ranks identical:", same, info_t["fun_evals"], info_1["fun_evals"], flush=True)
        ok = ok and same and et[0] < 1e-7 and et[1] < 1e-5 and info_t["fun_evals"] == info_1["fun_evals"]
        e = [abs(f - f1) / abs(f1), np.max(np.abs(g - g1)) / np.max(np.abs(g1)), np.max(np.abs(w - w1)) / np.max(np.abs(w1)),
             abs(nl[0, 0] - nl1[0, 0]) / abs(nl1[0, 0])] + [abs(st[k] - st1[k]) for k in st]
        print(method, "sharded vs single:", e, flush=True)
        ok = ok and max(e) < 1e-10
    dist.barrier()
if rank == 0:
    print("MGPU_OK" if ok else "MGPU_FAIL", flush=True)
dist.destroy_process_group()
sys.exit(0 if ok else 1)
