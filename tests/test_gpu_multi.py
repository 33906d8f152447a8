"""Two ranks on two GPUs over NCCL (skipped on a single-GPU box): row shards + ONE all-reduce of the packed buffer
through gpfy_allreduce_packed must reproduce the single-GPU result of the full data set to ~1e-13."""
import os

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, N, ret):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    import torch.distributed as dist
    from gpfy_b200 import distributed as gd
    from gpfy_b200.synthetic import headline_problem
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        full = headline_problem(N, T=12, F=6, device=dev, rank=0, world=1)      # every rank builds the same full data set
        lo, hi = gd.shard_bounds(N, rank, world)
        hp, args = full["hot"], full["args"]
        ref = hp.elbo_valgrad_packed(*args).clone()
        mine = hp.elbo_valgrad_packed(args[0][lo:hi].contiguous(), args[1][lo:hi].contiguous(), *args[2:]).clone()
        ar = gd.PackedAllReduce()
        assert ar.backend == "nccl" and ar.world == world and ar.comm is not None
        ar(mine)
        torch.cuda.synchronize()
        err = float((mine - ref).abs().max() / ref.abs().max())
        ar.close()
        ret[rank] = err
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_two_gpu_allreduce_matches_single_gpu():
    world, N = 2, 100_003
    ctx = mp.get_context("spawn")
    ret = ctx.Manager().dict()
    port = 29700 + os.getpid() % 200
    procs = [ctx.Process(target=_worker, args=(r, world, port, N, ret)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(300)
        assert p.exitcode == 0
    assert all(ret.get(r, 1.0) < 1e-12 for r in range(world)), dict(ret)


def _train_worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    import torch.distributed as dist
    from gpfy_b200 import distributed as gd
    from gpfy_b200.device_step import DeviceTrainer
    from gpfy_b200.likelihoods import Gaussian
    from gpfy_b200.model import GP
    from gpfy_b200.param import tree_leaves
    from gpfy_b200.spherical import PolynomialDecay
    from gpfy_b200.spherical_harmonics import SphericalHarmonics
    from gpfy_b200.variational import VariationalDistributionTriL
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        N, d = 50_001, 5
        g = torch.Generator(device=dev); g.manual_seed(7)
        X = torch.randn((N, d), dtype=torch.float64, device=dev, generator=g)
        Y = torch.sin(X[:, :1]) + 0.1 * torch.randn((N, 1), dtype=torch.float64, device=dev, generator=g)
        sh, q, lik = SphericalHarmonics(5, phase_truncation=12), VariationalDistributionTriL(), Gaussian()
        m = GP(PolynomialDecay()).conditional(sh, q)
        param = m.init(0, input_dim=d, likelihood=lik, sh_features=sh, variational_dist=q)
        single = DeviceTrainer(param, m, q, lik, learning_rate=2e-2)
        lo, hi = gd.shard_bounds(N, rank, world)
        sharded = DeviceTrainer(param, m, q, lik, learning_rate=2e-2, allreduce=gd.PackedAllReduce(), n_total=N)
        Xs, Ys = X[lo:hi].contiguous(), Y[lo:hi].contiguous()
        for _ in range(5):
            single.step(X, Y)
            sharded.step(Xs, Ys)
        err = max(float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))
                  for a, b in zip(tree_leaves(sharded.param().params), tree_leaves(single.param().params)))
        lerr = float(np.max(np.abs(sharded.losses() - single.losses()) / np.abs(single.losses())))
        # minibatches: every rank draws `batch` rows of its own shard; one rank stepping on the concatenation of both
        # ranks' batches with dataset_size = N must follow the same trajectory (optimization.py:154 rescale)
        from gpfy_b200.minibatch import DeviceMinibatches
        batch, steps = 512, 4
        mb_sharded = DeviceTrainer(param, m, q, lik, learning_rate=2e-2, allreduce=gd.PackedAllReduce(), n_total=N)
        mb_sharded.fit_minibatch(Xs, Ys, batch, steps, seed=11)
        mb_single = DeviceTrainer(param, m, q, lik, learning_rate=2e-2, dataset_size=N)
        gens = []
        for r in range(world):
            l2, h2 = gd.shard_bounds(N, r, world)
            gens.append(DeviceMinibatches(X[l2:h2].contiguous(), Y[l2:h2].contiguous(), batch, seed=11))
        for _ in range(steps):
            parts = [gen.next() for gen in gens]
            mb_single.step(torch.cat([p[0] for p in parts]), torch.cat([p[1] for p in parts]))
        merr = max(float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))
                   for a, b in zip(tree_leaves(mb_sharded.param().params), tree_leaves(mb_single.param().params)))
        mlerr = float(np.max(np.abs(mb_sharded.losses() - mb_single.losses()) / np.abs(mb_single.losses())))
        ret[rank] = max(err, lerr, merr, mlerr)
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_two_gpu_device_trainer_matches_single_gpu_training():
    world = 2
    ctx = mp.get_context("spawn")
    ret = ctx.Manager().dict()
    port = 29900 + os.getpid() % 90
    procs = [ctx.Process(target=_train_worker, args=(r, world, port, ret)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(300)
        assert p.exitcode == 0
    assert all(ret.get(r, 1.0) < 1e-9 for r in range(world)), dict(ret)


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_one_process_two_threads_two_devices():
    """XLA's execution model: ONE process, one host thread per device, concurrent calls into the library.
    The per-device state (SM count, shared-memory opt-in of the non-templated kernels) is keyed by device ordinal under
    a mutex, so the second device must not launch with the first one's configuration.  Each thread computes its half of
    the rows on its own device; the host sum equals the single-device result."""
    import threading
    from gpfy_b200.synthetic import headline_problem
    N = 150_003
    res, err = [None, None], [None, None]

    def work(i):
        try:
            torch.cuda.set_device(i)
            p = headline_problem(N, rank=i, world=2, device=torch.device("cuda", i))
            for _ in range(3):
                out = p["hot"].elbo_valgrad_packed(*p["args"])
                fmu, fvar = p["hot"].predict_diag(*[p["args"][k] for k in (0, 2, 3, 4, 5, 6, 7, 8, 9)])
            torch.cuda.synchronize(i)
            res[i] = (out.cpu(), fmu.cpu(), fvar.cpu())
        except Exception as e:  # noqa: BLE001
            err[i] = repr(e)

    th = [threading.Thread(target=work, args=(i,)) for i in (1, 0)]   # device 1 first: it must configure itself
    for t in th:
        t.start()
    for t in th:
        t.join()
    assert err == [None, None], err
    torch.cuda.set_device(0)
    full = headline_problem(N, rank=0, world=1, device=torch.device("cuda", 0))
    ref = full["hot"].elbo_valgrad_packed(*full["args"]).cpu()
    rmu, rvar = full["hot"].predict_diag(*[full["args"][k] for k in (0, 2, 3, 4, 5, 6, 7, 8, 9)])
    assert float((res[0][0] + res[1][0] - ref).abs().max() / ref.abs().max()) < 1e-12
    assert torch.allclose(torch.cat([res[0][1], res[1][1]]), rmu.cpu(), rtol=1e-12, atol=1e-13)
    assert torch.allclose(torch.cat([res[0][2], res[1][2]]), rvar.cpu(), rtol=1e-12, atol=1e-13)
