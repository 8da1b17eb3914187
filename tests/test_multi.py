"""Row-sharded evaluation: world_size-2 gloo test on CPU (host logic + the two-allreduce decomposition,
checked against the unsharded oracle), and a 2-GPU NCCL test of the CUDA path (skipped on <2 GPUs)."""
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _gloo_worker(rank, world, port, method, q):
    import torch
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import sharded_oracle as S
    from gpz_b200 import synth
    from oracle import gpz_oracle as O
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n, d, m, k = 301, 3, 9, 2
    X, Y = synth.make_data(n, d, seed=4, k=k)
    X, Y = np.array(X), np.array(Y)
    theta = synth.perturb_theta(synth.make_theta0(X, Y, method, m, het=True, seed=5), 0.1, 6)
    omega = 0.5 + np.random.default_rng(7).random((n, 1))
    model = O.Model(d=d, k=k, m=m, method=method, heteroscedastic=True)
    lo, hi = S.shard_bounds(n, rank, world)
    loc, p1 = S.sweep1(theta, model, X[lo:hi], Y[lo:hi], omega[lo:hi])
    t = torch.from_numpy(p1)
    dist.all_reduce(t)                                   # allreduce #1: Gram, rhs, scalars
    sol = S.solve(theta, model, t.numpy())
    p2 = S.sweep2(theta, model, X[lo:hi], Y[lo:hi], omega[lo:hi], loc, sol)
    t2 = torch.from_numpy(p2)
    dist.all_reduce(t2)                                  # allreduce #2: gradient partials, statistics
    f, g, st = S.assemble(theta, model, sol, t2.numpy())
    # the 128-byte communicator id travels over the same host channel in the product (bench.py)
    obj = [bytes(range(128)) if rank == 0 else None]
    dist.broadcast_object_list(obj, src=0)
    if rank == 0:
        ref = O.GPz(theta, model, X, Y, None, omega)
        q.put((f, g, st, ref.nlogML, ref.grad, ref.stats, obj[0] == bytes(range(128))))
    else:
        q.put(("uid", obj[0] == bytes(range(128))))
    dist.destroy_process_group()


@pytest.mark.parametrize("method", ["VD", "GL"])
def test_sharded_two_allreduce_decomposition_gloo(method):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000) + (0 if method == "VD" else 1)
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, method, q)) for r in range(2)]
    for p in procs:
        p.start()
    outs = [q.get(timeout=300) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    main = [o for o in outs if o[0] != "uid"][0]
    other = [o for o in outs if o[0] == "uid"][0]
    f, g, st, fr, gr, str_, uid_ok = main
    assert uid_ok and other[1]
    assert abs(f - fr) <= 1e-12 * abs(fr)
    assert np.max(np.abs(g - gr)) <= 1e-11 * np.max(np.abs(gr))
    assert abs(st["trainRMSE"] - str_["trainRMSE"]) <= 1e-12 and abs(st["trainLL"] - str_["trainLL"]) <= 1e-12


def _gloo_train_worker(rank, world, port, q):
    """The optimiser replicated on every rank over the sharded objective: the control flow needs no extra exchange
    because f and g are identical on all ranks after the two all-reduces."""
    import torch
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import sharded_oracle as S
    from gpz_b200 import synth
    from oracle import gpz_oracle as O
    from oracle import minfunc_oracle as MO
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n, d, m, k = 240, 2, 7, 1
    X, Y = synth.make_data(n, d, seed=8, k=k)
    X, Y = np.array(X), np.array(Y)
    theta0 = synth.make_theta0(X, Y, "VD", m, het=True, seed=9)
    omega = np.ones((n, 1))
    model = O.Model(d=d, k=k, m=m, method="VD", heteroscedastic=True)
    lo, hi = S.shard_bounds(n, rank, world)

    def fun_stats(theta):
        loc, p1 = S.sweep1(theta, model, X[lo:hi], Y[lo:hi], omega[lo:hi])
        t = torch.from_numpy(p1)
        dist.all_reduce(t)
        sol = S.solve(theta, model, t.numpy())
        t2 = torch.from_numpy(S.sweep2(theta, model, X[lo:hi], Y[lo:hi], omega[lo:hi], loc, sol))
        dist.all_reduce(t2)
        f, g, st = S.assemble(theta, model, sol, t2.numpy())
        return f, g, (st["trainRMSE"], st["trainLL"], np.nan, np.nan)

    x, best, bv, flag, info = MO.train_loop(fun_stats, theta0, theta0, -np.inf, max_iter=6, training_only=True)
    out = dict(rank=rank, x=x, bv=bv, evals=info["funcCount"])
    if rank == 0:
        def ref_stats(theta):
            r = O.GPz(theta, model, X, Y, None, omega)
            return r.nlogML, r.grad, (r.stats["trainRMSE"], r.stats["trainLL"], np.nan, np.nan)
        xr, _, bvr, _, infor = MO.train_loop(ref_stats, theta0, theta0, -np.inf, max_iter=6, training_only=True)
        out.update(xr=xr, bvr=bvr, evals_r=infor["funcCount"])
    q.put(out)
    dist.destroy_process_group()


def test_sharded_training_loop_gloo():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_gloo_train_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    outs = sorted([q.get(timeout=300) for _ in range(2)], key=lambda o: o["rank"])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    a, b = outs
    assert np.array_equal(a["x"], b["x"]) and a["bv"] == b["bv"] and a["evals"] == b["evals"]      # replicas agree bit for bit
    assert a["evals"] == a["evals_r"] and abs(a["bv"] - a["bvr"]) <= 1e-9 * abs(a["bvr"])
    assert np.max(np.abs(a["x"] - a["xr"])) <= 1e-7 * np.max(np.abs(a["xr"]))


def test_shard_bounds_partition():
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import sharded_oracle as S
    for n in (0, 1, 7, 1000, 10**6 + 3):
        for world in (1, 2, 4, 8):
            b = [S.shard_bounds(n, r, world) for r in range(world)]
            assert b[0][0] == 0 and b[-1][1] == n
            assert all(b[i][1] == b[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in b]
            assert max(sizes) - min(sizes) <= 1


@pytest.mark.gpu
def test_two_gpu_nccl_matches_single_gpu():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (run under gpurun --gpus 2)")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", str(29700 + os.getpid() % 200), os.path.join(ROOT, "tests", "mgpu_worker.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert "MGPU_OK" in out.stdout
