"""The ELBO and its training step (mirror of gpfy/optimization.py:60-156).

`elbo(param, m, q, lik, (X, Y), dataset_size=-1)` keeps the reference's signature.  The reference gets
gradients from `jax.value_and_grad` through the whole graph; here the N-dependent value AND all its
gradients come from ONE fused device call (`gpfy_svgp_elbo_valgrad`, sums over the local rows), are
all-reduced once when rows are sharded over ranks, and the N-independent chain rules (lambda_l(beta),
V -> Vhat -> chol Gram, bijectors, KL) are applied on the host — O(M^2 + sum m_l^3) numpy."""
import numpy as np
import torch

from . import distributed as gd
from .gegenbauer import gegenbauer_grad_host
from .param import Param, tree_map


def _packed(param, m, q, lik, train_data, allreduce):
    X, Y = train_data
    hp, args = m.hot_args(param, lik)
    lcat = args[5]
    s2 = lik._sigma2(param)
    if torch.is_tensor(X) and not X.is_cuda and X.is_pinned():
        Yh = Y if torch.is_tensor(Y) else torch.as_tensor(np.asarray(Y, dtype=np.float64))
        host = hp.elbo_valgrad_host(X, Yh, *args, s2, reduce=allreduce if (allreduce and allreduce.world > 1 and allreduce.backend == "nccl") else None)
        summed = host.numpy().copy()
        if allreduce is not None and allreduce.world > 1 and allreduce.backend != "nccl":
            summed = allreduce(torch.from_numpy(summed)).numpy()
    else:
        packed = hp.elbo_valgrad_packed(X, Y, *args, s2)
        if allreduce is not None and allreduce.world > 1:
            if allreduce.backend == "nccl":
                allreduce(packed)
                summed = packed.cpu().numpy()
            else:
                summed = allreduce(packed.cpu()).numpy()
        else:
            summed = packed.cpu().numpy()
    return hp, summed, (lcat if torch.is_tensor(lcat) else None)


def _chol_vjp(L, dL):
    """Cotangent of B given the cotangent of L = chol(B) (lower), symmetrised like jnp.linalg.cholesky."""
    P = np.tril(L.T @ np.tril(dL))
    P[np.diag_indices_from(P)] *= 0.5
    X = np.linalg.solve(L.T, P)                 # L^-T P   (scipy's solve_triangular costs ~5 ms per 30 x 30 call here)
    S = np.linalg.solve(L.T, X.T).T             # L^-T P L^-1
    return 0.5 * (S + S.T)


def _inducing_chain(sh, param, dvhat, dlcat, device_basis=None):
    """d/dV_l from (dVhat_l, dL_l): V -> V / |V| -> c_l C_l(Vhat Vhat^T) + 1e-16 I -> chol
    (spherical_harmonics.py:209-217, 265-288; gegenbauer.py:73-81).  Fixed bases: dV = dVhat.
    `device_basis` = (HotPath, lcat on the device): the chain runs in `gpfy_inducing_basis_vjp`."""
    raw = [v for _, v in sorted(param.params["variational"]["inducing_features"].items())]
    names = sorted(param.params["variational"]["inducing_features"])
    alpha = param.constants["sphere"]["alpha"]
    learned = sh._learned(param)
    if learned and device_basis is not None:
        hp, lcat_dev = device_basis
        dV = hp.inducing_basis_vjp(np.concatenate(raw, 0), lcat_dev, dvhat, dlcat, normalise=True).cpu().numpy()
        out, o = {}, 0
        for name, V in zip(names, raw):
            out[name] = dV[o:o + V.shape[0]].copy()
            o += V.shape[0]
        return out
    out, o, lo = {}, 0, 0
    Ls = sh.Ls(param) if learned else None
    for n, (name, V) in enumerate(zip(names, raw)):
        ml = V.shape[0]
        dVh = dvhat[o:o + ml]
        if learned:
            nrm = np.sqrt(np.sum(np.square(V), axis=1, keepdims=True))
            Vh = V / nrm
            dL = dlcat[lo:lo + ml * ml].reshape(ml, ml)
            dB = _chol_vjp(Ls[n], dL)
            G = Vh @ Vh.T
            dG = dB * (alpha / (alpha + float(n))) * gegenbauer_grad_host(n, alpha, G)
            dVh = dVh + (dG + dG.T) @ Vh
            out[name] = (dVh - Vh * np.sum(Vh * dVh, axis=1, keepdims=True)) / nrm
        else:
            out[name] = dVh.copy()
        o += ml
        lo += ml * ml
    return out


def elbo_value_and_grad(param: Param, m, q, lik, train_data, dataset_size: int = -1, allreduce=None, n_total=None):
    """(ELBO, gradient w.r.t. every leaf of the CONSTRAINED `param.params`, same nested structure).

    `train_data` are this rank's rows; with rows sharded over ranks pass `allreduce`
    (distributed.PackedAllReduce) and the global row count `n_total`."""
    if not param._constrained:
        raise ValueError("elbo expects constrained parameters (call param.constrained())")
    hp, summed, lcat_dev = _packed(param, m, q, lik, train_data, allreduce)
    n_local = int(np.shape(train_data[0])[0])
    n_total = int(n_total) if n_total is not None else n_local
    kl, dkl_mu, dkl_sig = q.prior_KL_and_grad(param)
    sl = gd.layout_slices(hp.layout, hp.M, hp.P, hp.F, hp.input_dim, hp.dim, hp.lsize, hp.scalar_w, hp.nw)
    value, g = gd.finish_elbo(summed, n_total, dataset_size, kl, dkl_mu, dkl_sig, sl)
    kernel, sh = m.kernel, m._sh
    kp = param.params[kernel.name]
    grads = tree_map(lambda p: np.zeros_like(p), param.params)
    gk = grads[kernel.name]
    gk["weight_variances"] = g["w"].reshape(np.shape(kp["weight_variances"]))
    gk["bias_variance"] = g["b"].reshape(())
    gk["variance"] = g["v"].reshape(())
    if "beta" in kp:  # lambda_l(beta), spherical.py:563-587
        gk["beta"] = np.array(np.sum(g["eig"] * kernel._eigenvalues_and_dbeta(param, sh.levels)[1]))
    if lik._sigma2(param) is not None:
        grads["likelihood"][lik.name]["variance"] = g["sigma2"].reshape(())
    gv = grads["variational"]
    gv[q.name]["mu"] = g["mu"].reshape(hp.M, hp.P)
    gv[q.name]["Sigma"] = g["sigma"].reshape(hp.P, hp.M, hp.M)
    gv["inducing_features"] = _inducing_chain(sh, param, g["vhat"].reshape(hp.M, hp.dim), g["lcat"],
                                              device_basis=(hp, lcat_dev) if lcat_dev is not None else None)
    return value, grads


def elbo(param: Param, m, q, lik, train_data, dataset_size: int = -1) -> float:
    """optimization.py:125-156: scale * mean_n(var_exp_n) - KL."""
    return elbo_value_and_grad(param, m, q, lik, train_data, dataset_size)[0]


def loss_and_grad_unconstrained(free_param: Param, m, q, lik, train_data, dataset_size: int = -1, allreduce=None, n_total=None):
    """(-ELBO, d(-ELBO)/d unconstrained params): what `jax.value_and_grad(loss_fn)(state.params)` returns at
    optimization.py:113-118; the bijector VJPs of param.py:228-258 are applied here."""
    con = free_param.constrained()
    value, g = elbo_value_and_grad(con, m, q, lik, train_data, dataset_size, allreduce, n_total)
    gfree = tree_map(lambda x, gc, t: -t.forward_vjp(x, gc), free_param.params, g, free_param._bijectors)
    return -value, gfree


def _materialise_rows(X, Y, device=None):
    """(X, Y) as float64 CUDA tensors (used in place if they already are), or pinned host tensors when the rows do not
    fit in the free device memory next to the hot path's workspace (kept at <= 40 % of what is free now)."""
    from .ops import require_cuda
    dev = require_cuda(device)
    if torch.is_tensor(X) and (X.is_cuda or X.is_pinned()):
        Yt = Y if torch.is_tensor(Y) else torch.as_tensor(np.asarray(Y, dtype=np.float64))
        if X.is_cuda and not Yt.is_cuda:
            Yt = Yt.to(device=X.device, dtype=torch.float64)
        return X, Yt
    Xt = X if torch.is_tensor(X) else torch.from_numpy(np.ascontiguousarray(np.asarray(X, dtype=np.float64)))
    Yt = Y if torch.is_tensor(Y) else torch.from_numpy(np.ascontiguousarray(np.asarray(Y, dtype=np.float64)))
    Xt, Yt = Xt.to(torch.float64).contiguous(), Yt.to(torch.float64).reshape(Xt.shape[0], -1).contiguous()
    need = (Xt.numel() + Yt.numel()) * 8
    free, _ = torch.cuda.mem_get_info(dev)
    if need <= 0.4 * free:
        return Xt.to(dev), Yt.to(dev)
    return Xt.pin_memory(), Yt.pin_memory()


def create_training_step(model, dataset, dataset_xy_keys, q, lik, batch_size=None, allreduce=None, n_total=None, seed=0):
    """optimization.py:60-122.  `dataset` is any mapping of column name -> array (the reference takes a
    HuggingFace Dataset; `datasets` is a host-side container either way).  Full batch: the rows are captured
    once; minibatch: a fresh shuffled batch every step, gathered on the device by `gpfy_minibatch_gather`
    (gpfy_b200.minibatch), with the `dataset_size = N` rescale (the reference freezes its first batch at trace
    time — a quirk that is not replicated)."""
    x_key, y_key = dataset_xy_keys
    X_all, Y_all = dataset[x_key], dataset[y_key]
    N = int(np.shape(X_all)[0])
    world = allreduce.world if allreduce is not None else 1
    if world > 1 and n_total is None:
        raise ValueError("create_training_step: rows sharded over ranks (allreduce.world > 1) need n_total, the global row count")
    if batch_size:
        from .minibatch import DeviceMinibatches
        batches = DeviceMinibatches(X_all, Y_all, batch_size, seed=seed)
        next_batch = batches.next
        n_global = N if n_total is None else int(n_total)

        def train_step(state, _=None):
            X, Y = next_batch()
            # optimization.py:154: scale = dataset_size / batch; under data parallelism every rank draws `batch_size`
            # rows of its own shard, so the batch is batch_size * world rows of the n_total-row data set
            loss, grads = loss_and_grad_unconstrained(state.apply_fn(state.params), model, q, lik, (X, Y),
                                                      dataset_size=n_global, allreduce=allreduce, n_total=batch_size * world)
            return state.apply_gradients(grads=grads), loss
        return train_step

    # full batch: the rows are materialised ONCE - on the device when they fit next to the workspace, otherwise in
    # pinned host memory, from where every step streams them (HotPath.elbo_valgrad_host) - and reused by every step
    X_all, Y_all = _materialise_rows(X_all, Y_all)

    def train_step(state, _=None):
        loss, grads = loss_and_grad_unconstrained(state.apply_fn(state.params), model, q, lik, (X_all, Y_all),
                                                  allreduce=allreduce, n_total=n_total)
        return state.apply_gradients(grads=grads), loss
    return train_step
