"""Training step driver on the device.

The reference's `train_step` (optimization.py:107-120) applies the bijectors, evaluates the ELBO and its gradient,
and lets optax update the unconstrained parameters - all inside one jitted XLA program.  `DeviceTrainer` is the
equivalent around the CUDA hot path: the unconstrained parameters live in ONE flat device vector, and a step is a
fixed sequence of stream-ordered C-ABI calls (constrain -> inducing basis -> ELBO + gradient -> all-reduce -> KL ->
basis VJP -> gradient assembly -> Adam) without a single host synchronisation; `fit(..., use_graph=True)` /
`fit_minibatch(..., use_graph=True)` capture that sequence once in a CUDA graph and replay it.  Supports the TriL variational distribution with Gaussian or Bernoulli likelihood and every kernel of
gpfy_b200.spherical; m_l <= 256 in learned-basis mode.
"""
import ctypes

import numpy as np
import torch

from . import _lib
from .ops import _ptr, _stream
from .param import Exp, FillTriangular, Identity, Param, tree_map
from .spherical import PolynomialDecay


class DeviceTrainer:
    def __init__(self, param: Param, m, q, lik, learning_rate=5e-2, b1=0.9, b2=0.999, eps=1e-8, dataset_size=-1,
                 allreduce=None, n_total=None):
        if q._q_kind != "tril":
            raise _lib.GpfyError("DeviceTrainer supports VariationalDistributionTriL (the full-covariance KL needs a host Cholesky)")
        self.m, self.q, self.lik, self.kernel, self.sh = m, q, lik, m.kernel, m._sh
        self.lr, self.b1, self.b2, self.eps = float(learning_rate), float(b1), float(b2), float(eps)
        self.dataset_size, self.allreduce, self.n_total = dataset_size, allreduce, n_total
        self.template = param if param._constrained else param.constrained()
        con = self.template
        self.hp, _ = m.hot_args(con, lik)
        hp, dev = self.hp, self.hp.device
        kname = self.kernel.name
        kp = con.params[kname]
        self.learned = self.sh._learned(con)
        if self.learned and max(hp.m) > hp.BASIS_MAX_M:
            raise _lib.GpfyError(f"DeviceTrainer needs m_l <= {hp.BASIS_MAX_M} in learned-basis mode")
        self.vnames = sorted(con.params["variational"]["inducing_features"])
        self._check_bijectors(con, kname, lik, q, bool(hp.projection_dim))
        M, P, F, dim = hp.M, hp.P, hp.F, hp.dim
        ntri = M * (M + 1) // 2
        L = _lib.GpfyStepLayout()
        L.kernel_kind = 1 if isinstance(self.kernel, PolynomialDecay) else 0
        L.truncation_level = getattr(self.kernel, "truncation_level", 0)
        L.w_identity = 1 if hp.projection_dim else 0
        L.n_w = hp.nw
        o = 0
        def take(n):
            nonlocal o
            r = o
            o += n
            return r
        L.th_w, L.th_b, L.th_v = take(hp.nw), take(1), take(1)
        L.th_beta = take(1) if "beta" in kp else -1
        self.has_s2 = lik._sigma2(con) is not None
        L.th_s2 = take(1) if self.has_s2 else -1
        L.th_mu, L.th_sig, L.th_V = take(M * P), take(P * ntri), take(M * dim)
        L.th_total = o
        o = 0
        L.c_w, L.c_b, L.c_v, L.c_beta, L.c_eig, L.c_deig, L.c_s2 = take(hp.nw), take(1), take(1), take(1), take(F), take(F), take(1)
        L.c_mu, L.c_sig, L.c_V = take(M * P), take(P * M * M), take(M * dim)
        L.c_total = o
        self.L = L
        # flat unconstrained vector + trainable mask from the Param
        free = con.unconstrained()
        fk, tk = free.params[kname], con._trainables[kname]
        theta = np.zeros(L.th_total)
        mask = np.zeros(L.th_total, dtype=np.uint8)
        def put(off, val, tr):
            val = np.asarray(val, dtype=np.float64).ravel()
            theta[off:off + val.size] = val
            mask[off:off + val.size] = 1 if tr else 0
        put(L.th_w, fk["weight_variances"], tk["weight_variances"])
        put(L.th_b, fk["bias_variance"], tk["bias_variance"])
        put(L.th_v, fk["variance"], tk["variance"])
        if L.th_beta >= 0:
            put(L.th_beta, fk["beta"], tk["beta"])
        if self.has_s2:
            put(L.th_s2, free.params["likelihood"][lik.name]["variance"], con._trainables["likelihood"][lik.name]["variance"])
        fq, tq = free.params["variational"][q.name], con._trainables["variational"][q.name]
        put(L.th_mu, fq["mu"], tq["mu"])
        put(L.th_sig, fq["Sigma"], tq["Sigma"])
        off = L.th_V
        for name in self.vnames:
            v = free.params["variational"]["inducing_features"][name]
            put(off, v, con._trainables["variational"]["inducing_features"][name])
            off += v.size
        t = lambda a, dt=torch.float64: torch.as_tensor(np.ascontiguousarray(a), device=dev).to(dt)
        self.theta, self.mask = t(theta), torch.as_tensor(mask, device=dev)
        self.m1, self.m2 = torch.zeros_like(self.theta), torch.zeros_like(self.theta)
        self.grad = torch.zeros_like(self.theta)
        self.cons = torch.zeros(L.c_total, dtype=torch.float64, device=dev)
        flat, src = FillTriangular._index_map(M)
        tri = np.zeros(ntri, dtype=np.int32)
        tri[src] = flat
        self.tri_map = torch.as_tensor(tri, device=dev)
        if L.kernel_kind == 1:
            cf = self.kernel._compute_eigvals(con.constants["sphere"]["sphere_dim"],
                                              gegenbauer_lookup_table=con.constants["sphere"]["gegenbauer_lookup_table"])
        else:
            cf = self.kernel.eigenvalues(con, self.sh.levels)
        self.cf = t(cf)
        self.packed = torch.zeros(hp.layout.total, dtype=torch.float64, device=dev)
        self.kl = torch.zeros(1 + M * P + P * M * M, dtype=torch.float64, device=dev)
        self.dV = torch.zeros(M * dim, dtype=torch.float64, device=dev)
        if self.learned:
            self.vhat = torch.zeros((M, dim), dtype=torch.float64, device=dev)
            self.lcat = torch.zeros(hp.lsize, dtype=torch.float64, device=dev)
        else:
            self.lcat = t(np.concatenate([l.ravel() for l in self.sh.Ls(con)]))
        self.loss_trace = []
        self.count = 0
        c = self.cons
        sl = lambda a, n: c[a:a + n]
        self.v_w, self.v_b, self.v_v, self.v_eig = sl(L.c_w, hp.nw), sl(L.c_b, 1), sl(L.c_v, 1), sl(L.c_eig, F)
        self.v_s2, self.v_mu, self.v_sig, self.v_V = sl(L.c_s2, 1), sl(L.c_mu, M * P), sl(L.c_sig, P * M * M), sl(L.c_V, M * dim)

    def _check_bijectors(self, con, kname, lik, q, projection):
        """The device kernels implement ONE transform per slot (gpfy_step_constrain / gpfy_step_grad: Exp for the
        kernel scalars, weights and the noise variance, Identity for projection weights, mu and V, FillTriangular for
        Sigma - the defaults of the reference, param.py:24-25, variational.py:260).  A Param whose bijector was changed
        with `set_bijector` would train with the wrong transform, so it is refused here (use the host loop
        `GP.fit` + `create_training_step`, which honours `param._bijectors`)."""
        bj = con._bijectors
        want = [((kname, "weight_variances"), Identity if projection else Exp), ((kname, "bias_variance"), Exp),
                ((kname, "variance"), Exp)]
        if "beta" in con.params[kname]:
            want.append(((kname, "beta"), Exp))
        if lik._sigma2(con) is not None:
            want.append((("likelihood", lik.name, "variance"), Exp))
        want += [(("variational", q.name, "mu"), Identity), (("variational", q.name, "Sigma"), FillTriangular)]
        want += [(("variational", "inducing_features", n), Identity) for n in self.vnames]
        for path, cls in want:
            node = bj
            for k in path:
                node = node[k]
            if type(node) is not cls:
                raise _lib.GpfyError(f"DeviceTrainer: parameter {'/'.join(path)} has bijector {type(node).__name__}, the device step "
                                     f"implements {cls.__name__} there; train this Param with GP.fit + create_training_step instead")

    # ------------------------------------------------------------------
    def step(self, X, Y):
        """One Adam step on this rank's rows (device tensors); every call only enqueues work.  The loss (-ELBO) of the
        step stays on the device in `loss_trace`."""
        loss = torch.empty(1, dtype=torch.float64, device=self.hp.device)
        self._enqueue_gradient(X, Y, loss)
        hp, L, lib = self.hp, self.L, self.hp.lib
        self.count += 1
        _lib.check(lib.gpfy_step_adam(_stream(hp.device), L.th_total, _ptr(self.theta), _ptr(self.grad), _ptr(self.m1), _ptr(self.m2),
                                      _ptr(self.mask), self.lr, self.b1, self.b2, self.eps, self.count))
        self.loss_trace.append(loss)
        return loss

    def _enqueue_gradient(self, X, Y, loss, kl_stream=None):
        """constrain -> inducing basis -> ELBO + gradient -> all-reduce -> KL -> basis VJP -> gradient of -ELBO w.r.t. theta.
        `kl_stream`: run the two KL kernels on that stream, forked after `constrain` and joined before the gradient
        assembly (inside a graph capture this becomes a parallel branch; eagerly it would only add host work)."""
        hp, L, lib = self.hp, self.L, self.hp.lib
        st, d = _stream(hp.device), ctypes.byref(hp.desc)
        _lib.check(lib.gpfy_step_constrain(st, d, ctypes.byref(L), _ptr(self.theta), _ptr(self.tri_map), _ptr(self.cf), _ptr(self.cons)))
        if kl_stream is not None:
            main = torch.cuda.current_stream(hp.device)
            kl_stream.wait_stream(main)
            with torch.cuda.stream(kl_stream):
                _lib.check(lib.gpfy_prior_kl_valgrad(_stream(hp.device), d, _ptr(self.v_mu), _ptr(self.v_sig), _ptr(self.kl)))
        if self.learned:
            bws = hp._basis_workspace()
            _lib.check(lib.gpfy_inducing_basis_fwd_ws(st, d, _ptr(self.v_V), 1, _ptr(self.vhat), _ptr(self.lcat), _ptr(bws), bws.numel() * 8))
            vhat = self.vhat
        else:
            vhat = self.v_V
        # the sigma2 slot of `cons` is passed for both likelihoods (ignored for Bernoulli): no per-step host scalar
        hp.elbo_valgrad_packed(X, Y, self.v_w, self.v_b, self.v_v, self.v_eig, vhat, self.lcat, self.v_mu, self.v_sig,
                               self.v_s2, out=self.packed)
        n_local = X.shape[0]
        if self.allreduce is not None and self.allreduce.world > 1:
            self.allreduce(self.packed)
        n_total = int(self.n_total) if self.n_total is not None else n_local
        scale = float(self.dataset_size) if self.dataset_size > 0 else float(n_total)
        if kl_stream is None:
            _lib.check(lib.gpfy_prior_kl_valgrad(st, d, _ptr(self.v_mu), _ptr(self.v_sig), _ptr(self.kl)))
        E = hp.layout
        if self.learned:
            bws = hp._basis_workspace()
            _lib.check(lib.gpfy_inducing_basis_vjp_ws(st, d, _ptr(self.v_V), _ptr(self.lcat), 1,
                                                      ctypes.c_void_p(self.packed.data_ptr() + 8 * E.d_vhat),
                                                      ctypes.c_void_p(self.packed.data_ptr() + 8 * E.d_lcat), _ptr(self.dV), _ptr(bws), bws.numel() * 8))
            dV = self.dV
        else:
            dV = self.packed[E.d_vhat:E.d_vhat + hp.M * hp.dim]
        if kl_stream is not None:
            torch.cuda.current_stream(hp.device).wait_stream(kl_stream)
        _lib.check(lib.gpfy_step_grad(st, d, ctypes.byref(L), _ptr(self.cons), _ptr(self.packed), _ptr(self.kl), _ptr(dV),
                                      _ptr(self.tri_map), scale / n_total, _ptr(self.grad), _ptr(loss)))

    # ------------------------------------------------------------------
    # CUDA-graph replay: the step count, the epoch and the position inside it live in `state` on the device
    # (gpfy_step_adam_state / gpfy_minibatch_gather_state / gpfy_step_advance), so a step has no per-step host argument.
    def _enqueue_state_step(self, X, Y, batches, state, loss_cur, trace, kl_stream=None):
        hp, L, lib = self.hp, self.L, self.hp.lib
        if batches is not None:
            X, Y = batches.next_from_state(state)
        self._enqueue_gradient(X, Y, loss_cur, kl_stream)
        st = _stream(hp.device)
        _lib.check(lib.gpfy_step_adam_state(st, L.th_total, _ptr(self.theta), _ptr(self.grad), _ptr(self.m1), _ptr(self.m2),
                                            _ptr(self.mask), self.lr, self.b1, self.b2, self.eps, _ptr(state)))
        _lib.check(lib.gpfy_step_advance(st, _ptr(state), batches.n if batches is not None else 0,
                                         batches.batch_size if batches is not None else 0, _ptr(loss_cur), _ptr(trace), trace.numel()))

    def _run_graph(self, X, Y, batches, num_steps):
        """`num_steps` steps: one eager step (sizes the scratch), then one captured step replayed num_steps - 1 times."""
        if self.allreduce is not None and self.allreduce.world > 1:
            raise _lib.GpfyError("use_graph is single-rank (the all-reduce is enqueued outside graph capture)")
        dev = self.hp.device
        num_steps = int(num_steps)
        if num_steps <= 0:
            return self
        if batches is None:
            X, Y = self.hp._dev(X), self.hp._dev(Y)
        state = torch.tensor([self.count, batches.epoch if batches is not None else 0, batches.pos if batches is not None else 0],
                             dtype=torch.int64, device=dev)
        trace = torch.zeros(num_steps, dtype=torch.float64, device=dev)
        loss_cur = torch.zeros(1, dtype=torch.float64, device=dev)
        # trace slots are indexed by the absolute step count modulo the capacity; rotate back afterwards
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            self._enqueue_state_step(X, Y, batches, state, loss_cur, trace)
            if num_steps > 1:
                graph = torch.cuda.CUDAGraph()
                kl_stream = torch.cuda.Stream(device=dev)
                with torch.cuda.graph(graph, stream=side):
                    self._enqueue_state_step(X, Y, batches, state, loss_cur, trace, kl_stream)
                for _ in range(num_steps - 1):     # capture only records; every further step is a replay
                    graph.replay()
        torch.cuda.current_stream(dev).wait_stream(side)
        start = self.count % num_steps
        self.loss_trace.append(torch.roll(trace, -start) if start else trace)
        self.count += num_steps
        if batches is not None:
            st = state.cpu().numpy()
            batches.epoch, batches.pos = int(st[1]), int(st[2])
        return self

    def fit(self, X, Y, num_steps, use_graph=True):
        """`num_steps` full-batch Adam steps on device tensors (model.py:261-302 without the host in the loop)."""
        if use_graph:
            return self._run_graph(X, Y, None, num_steps)
        for _ in range(int(num_steps)):
            self.step(X, Y)
        return self

    def fit_minibatch(self, X, Y, batch_size, num_steps, seed=0, use_graph=False):
        """`num_steps` Adam steps on shuffled minibatches of (X, Y) with the `dataset_size = N` rescale
        (optimization.py:84-104, 154): gather and step are both stream-ordered, nothing returns to the host.  Under
        data parallelism every rank samples `batch_size` rows of its own shard.  `use_graph`: capture one step (gather
        included) in a CUDA graph and replay it - for small batches the step is launch-bound and this removes the
        per-launch host cost."""
        from .minibatch import DeviceMinibatches
        batches = DeviceMinibatches(X, Y, batch_size, seed=seed, device=self.hp.device)
        world = self.allreduce.world if self.allreduce is not None else 1
        saved = self.dataset_size, self.n_total
        # the global row count (the constructor's n_total under data parallelism; shards may differ by one row)
        n_global = int(self.n_total) if self.n_total is not None else batches.n * world
        self.dataset_size, self.n_total = n_global, batch_size * world
        try:
            if use_graph:
                self._run_graph(None, None, batches, num_steps)
            else:
                for _ in range(int(num_steps)):
                    Xb, Yb = batches.next()
                    self.step(Xb, Yb)
        finally:
            self.dataset_size, self.n_total = saved
        return self

    def losses(self):
        return torch.cat(self.loss_trace).cpu().numpy() if self.loss_trace else np.zeros(0)

    def param(self) -> Param:
        """The current parameters as a constrained `Param` (one device -> host copy)."""
        L, hp = self.L, self.hp
        th = self.theta.cpu().numpy()
        free = self.template.unconstrained()
        p = tree_map(lambda x: x, free.params)
        kname = self.kernel.name
        kp = dict(p[kname])
        kp["weight_variances"] = th[L.th_w:L.th_w + hp.nw].reshape(np.shape(kp["weight_variances"]))
        kp["bias_variance"], kp["variance"] = th[L.th_b].reshape(()), th[L.th_v].reshape(())
        if L.th_beta >= 0:
            kp["beta"] = th[L.th_beta].reshape(())
        p[kname] = kp
        if self.has_s2:
            p["likelihood"] = {self.lik.name: {"variance": th[L.th_s2].reshape(())}}
        var = {k: (dict(v) if isinstance(v, dict) else v) for k, v in p["variational"].items()}
        M, P = hp.M, hp.P
        ntri = M * (M + 1) // 2
        var[self.q.name] = {"mu": th[L.th_mu:L.th_mu + M * P].reshape(M, P), "Sigma": th[L.th_sig:L.th_sig + P * ntri].reshape(P, ntri)}
        feats, off = {}, L.th_V
        for name in self.vnames:
            shp = np.shape(free.params["variational"]["inducing_features"][name])
            n = int(np.prod(shp))
            feats[name] = th[off:off + n].reshape(shp)
            off += n
        var["inducing_features"] = feats
        p["variational"] = var
        return free.replace(params=p).constrained()
