"""world_size-2 gloo test of the N>1 host logic: contiguous row shards, one all-reduce of
the packed [data term, gradients] buffer, then scale / N and the (never reduced) KL.  The per-shard packed
buffers come from the oracle (CPU); the GPU kernels are covered by the `-m gpu` tests."""
import ctypes
import os

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

from gpfy_b200 import _lib, distributed as gd
from oracle import gpfy_oracle as O
from tests.golden_util import golden_pack, load, rel_err


def test_shard_bounds_cover_rows_once():
    for n in (0, 1, 7, 64, 10_000_000):
        for world in (1, 2, 3, 8):
            b = [gd.shard_bounds(n, r, world) for r in range(world)]
            assert b[0][0] == 0 and b[-1][1] == n
            assert all(b[i][1] == b[i + 1][0] for i in range(world - 1))
            assert max(h - l for l, h in b) - min(h - l for l, h in b) <= 1
    with pytest.raises(ValueError):
        gd.shard_bounds(10, 2, 2)


def _resolved(name):
    g = load(name)
    pack = golden_pack(g)
    Vh, L = O.resolve_V_L(O.pack_to_torch(pack))
    pack["V"], pack["L"] = [x.numpy() for x in Vh], [x.numpy() for x in L]
    return g, pack


def _oracle_packed(pack, X, Y, sl, total):
    val, og = O.elbo_and_grads(pack, X, Y, data_only=True)
    buf = np.zeros(total)
    buf[sl["data_term"]] = val
    buf[sl["mu"]] = og["mu"].ravel()
    buf[sl["sigma"]] = og["Lq"].ravel()
    buf[sl["eig"]] = og["eig"]
    buf[sl["vhat"]] = np.concatenate([og[f"V{i}"] for i in range(len(pack["V"]))], 0).ravel()
    buf[sl["lcat"]] = np.concatenate([np.tril(og[f"L{i}"]).ravel() for i in range(len(pack["L"]))])
    buf[sl["w"]] = np.atleast_1d(og["w"])
    buf[sl["b"]], buf[sl["v"]], buf[sl["sigma2"]] = og["b"], og["v"], og["sigma2"]
    return buf


def _worker(rank, world, port, name, ret):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    import torch.distributed as dist
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g, pack = _resolved(name)
        X, Y = g["X"], g["Y"]
        N = X.shape[0]
        m = [v.shape[0] for v in pack["V"]]
        M, P, F, d = sum(m), pack["mu"].shape[1], len(m), X.shape[1]
        desc = _lib.make_desc(d, m, num_outputs=P)
        lay = _lib.GpfyElboLayout()
        _lib.check(_lib.load().gpfy_elbo_packed_layout(ctypes.byref(desc), ctypes.byref(lay)))
        sl = gd.layout_slices(lay, M, P, F, d, d + 1, sum(v * v for v in m), False)
        lo, hi = gd.shard_bounds(N, rank, world)
        packed = torch.from_numpy(_oracle_packed(pack, X[lo:hi], Y[lo:hi], sl, lay.total))
        ar = gd.PackedAllReduce()
        assert ar.world == world and ar.backend == "gloo"
        ar(packed)
        mu_t, L_t = O.T(pack["mu"]).requires_grad_(True), O.T(pack["Lq"]).requires_grad_(True)
        kl = O.prior_KL(mu_t, L_t, tril=True)
        kl.backward()
        value, grads = gd.finish_elbo(packed.numpy(), N, -1, float(kl), mu_t.grad.numpy(), L_t.grad.numpy(), sl)
        # the reference objective on ALL rows (optimization.py:151-156)
        full, fg = O.elbo_and_grads(pack, X, Y)
        ok = abs(value - full) <= 1e-12 * abs(full)
        ok &= rel_err(grads["mu"], fg["mu"].ravel()) < 1e-11 and rel_err(grads["sigma"], fg["Lq"].ravel()) < 1e-11
        ok &= rel_err(grads["w"], np.atleast_1d(fg["w"])) < 1e-11 and rel_err(grads["sigma2"], np.atleast_1d(fg["sigma2"])) < 1e-11
        ok &= rel_err(grads["vhat"], np.concatenate([fg[f"V{i}"] for i in range(F)], 0).ravel()) < 1e-11
        # minibatch rescale (optimization.py:154): dataset_size > 0 multiplies the data term only
        v2, _ = gd.finish_elbo(packed.numpy(), N, 5 * N, float(kl), mu_t.grad.numpy(), L_t.grad.numpy(), sl)
        ok &= abs(v2 - (5 * (full + float(kl)) - float(kl))) <= 1e-11 * abs(v2)
        ret[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("name", ["d8_F6_T12"])
def test_two_rank_allreduce_matches_full_batch_oracle(name):
    world = 2
    ctx = mp.get_context("spawn")
    ret = ctx.Manager().dict()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, world, port, name, ret)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(180)
        assert p.exitcode == 0
    assert all(ret.get(r) for r in range(world)), dict(ret)
