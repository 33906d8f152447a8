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
