"""GPU parity of the remaining C-ABI entry points: to_sphere, variational_expectations, project_q against
the reference goldens; the begin/rows/finish stages and the host-streamed path against the one-shot call;
full-size invariants at BASELINE.json's shapes (SURVEY.md §8c ii)."""
import ctypes
import math

import numpy as np
import pytest
import torch

from oracle import gpfy_oracle as O
from tests.golden_util import CASES, elem_err, golden_pack, load, rel_err
from tests.test_gpu_parity import ARD_CASES, _cat, _hot, _resolved

pytestmark = pytest.mark.gpu
TOL = 1e-10


@pytest.mark.parametrize("name", ARD_CASES)
def test_to_sphere_matches_reference_golden(name):
    g, pack = _resolved(name)
    hp = _hot(pack, g)
    xs, r = hp.to_sphere(g["X"], pack["w"], pack["b"])
    assert rel_err(xs.cpu().numpy(), g["o_x_sphere"]) < 1e-14
    assert rel_err(r.cpu().numpy(), g["o_r"]) < 1e-14
    # tests/test_spherical.py:103-120 invariants: unit norm, |scaled|^2 = r^2 - b
    assert float((xs.norm(dim=1) - 1).abs().max()) < 1e-14


@pytest.mark.parametrize("name", ARD_CASES)
def test_variational_expectations_match_reference_golden(name):
    g, pack = _resolved(name)
    hp = _hot(pack, g)
    ve = hp.variational_expectations(g["o_fmu"], g["o_fvar"], g["Y"], pack.get("sigma2"))
    assert rel_err(ve.cpu().numpy(), g["o_var_exp"]) < TOL
    # tests/test_likelihoods.py:32-46: var_exp(f, 0) == log_prob(f)
    f = torch.as_tensor(g["o_fmu"])
    y = torch.as_tensor(g["Y"])
    ve0 = hp.variational_expectations(g["o_fmu"], np.zeros_like(g["o_fvar"]), g["Y"], pack.get("sigma2")).cpu()
    if pack["lik"] == "gaussian":
        lp = O.gaussian_var_exp(O.T(pack["sigma2"]), f, torch.zeros_like(f), y)
    else:
        lp = O.bernoulli_log_prob(f, y)
    assert rel_err(ve0.numpy(), lp.numpy()) < TOL


@pytest.mark.parametrize("name", ARD_CASES)
def test_project_q_matches_reference_golden(name):
    g, pack = _resolved(name)
    hp = _hot(pack, g)
    sigma = pack.get("Lq", pack.get("Sigma"))
    mean, var = hp.project_q(g["o_proj"], pack["mu"], sigma)
    assert rel_err(mean.cpu().numpy(), g["o_project_mean"]) < TOL
    assert rel_err(var.cpu().numpy(), g["o_project_diag_variance"]) < TOL


def _problem(N, T=30, P=1, lik="gaussian", d=8, F=11):
    from gpfy_b200.synthetic import headline_problem
    return headline_problem(N, d=d, F=F, T=T, P=P, likelihood=lik, device="cuda")


@pytest.mark.parametrize("N", [1, 31, 1000, 150_001])
def test_stages_and_host_stream_equal_one_shot(N):
    """begin / rows (ragged slabs, empty slab) / finish == one call; the pinned-host streamed path too."""
    from gpfy_b200 import _lib
    from gpfy_b200.ops import _ptr, _stream
    prob = _problem(N)
    hp, args = prob["hot"], prob["args"]
    ref = hp.elbo_valgrad_packed(*args).clone()
    X, Y, w, b, v, eig, vh, lc, mu, sg, s2 = args
    cap = max(1, (N + 2) // 3)
    ws = hp.workspace(cap, _lib.OP_ELBO)
    d = ctypes.byref(hp.desc)
    out = torch.empty_like(ref)
    st = _stream(hp.device)
    _lib.check(hp.lib.gpfy_svgp_elbo_begin(st, d, cap, _ptr(w), _ptr(b), _ptr(v), _ptr(eig), _ptr(vh), _ptr(lc), _ptr(mu), _ptr(sg),
                                           _ptr(ws), ws.numel()))
    bounds = [0, min(N, cap // 2), min(N, cap // 2), min(N, cap // 2 + cap)]
    while bounds[-1] < N:
        bounds.append(min(N, bounds[-1] + cap))
    for lo, hi in zip(bounds[:-1], bounds[1:]):
        xs, ys = X[lo:hi].contiguous(), Y[lo:hi].contiguous()
        _lib.check(hp.lib.gpfy_svgp_elbo_rows(st, d, cap, hi - lo, _ptr(xs), _ptr(ys), _ptr(v), _ptr(s2), _ptr(ws), ws.numel()))
    _lib.check(hp.lib.gpfy_svgp_elbo_finish(st, d, cap, _ptr(w), _ptr(b), _ptr(v), _ptr(eig), _ptr(mu), _ptr(out), _ptr(ws), ws.numel()))
    torch.cuda.synchronize()
    a, r = hp.unpack(out), hp.unpack(ref)
    for k in r:
        assert rel_err(a[k].cpu().numpy(), r[k].cpu().numpy()) < 1e-11, k
    host = hp.elbo_valgrad_host(X.cpu().pin_memory(), Y.cpu().pin_memory(), *args[2:])
    h = hp.unpack(host)
    for k in r:
        assert rel_err(h[k].numpy(), r[k].cpu().numpy()) < 1e-11, k


@pytest.mark.parametrize("lik,P,T,F,d", [("gaussian", 1, 30, 11, 8), ("bernoulli", 1, 30, 11, 8), ("gaussian", 2, 12, 6, 8),
                                          ("gaussian", 1, 30, 13, 8), ("bernoulli", 1, 24, 5, 32)])
def test_several_internal_chunks_match_oracle(lik, P, T, F, d):
    """Rows spanning 3.2 internal chunks (chunk = 192 * SMs rows with the development switch chunk_mult = 1), a q far from
    the prior (random mu, L_q with O(1) off-diagonal mass): value and EVERY gradient against the oracle on all rows,
    1e-10 normwise and 1e-9 componentwise.  Covers the accumulation across chunks of every split-K / partial-sum buffer
    on the fused (P = 1, Gaussian), fused-backward (Bernoulli) and general (P = 2) paths."""
    from gpfy_b200 import ops
    from gpfy_b200.synthetic import headline_problem
    sms = torch.cuda.get_device_properties(0).multi_processor_count
    N = int(3.2 * 192 * sms) + 5
    prob = headline_problem(N, d=d, F=F, T=T, P=P, likelihood=lik, device="cuda", seed=11)
    hp, args, host = prob["hot"], list(prob["args"]), prob["host"]
    rng = np.random.default_rng(5)
    M = hp.M
    host["mu"] = rng.standard_normal((M, P))
    host["Lq"] = np.stack([np.tril(0.6 * np.eye(M) + 0.8 * rng.standard_normal((M, M)) / np.sqrt(M)) for _ in range(P)])
    args[8] = torch.as_tensor(host["mu"], device="cuda")
    args[9] = torch.as_tensor(host["Lq"], device="cuda")
    try:
        ops.debug_option("chunk_mult", 1)
        rows = ctypes.c_int64(0)
        hp.lib.gpfy_chunk_rows(ctypes.byref(hp.desc), N, ctypes.byref(rows))
        assert rows.value == 192 * sms and N > 3 * rows.value     # really four launches of every per-chunk kernel
        out = {k: v.cpu().numpy() for k, v in hp.unpack(hp.elbo_valgrad_packed(*args)).items()}
    finally:
        ops.debug_option("chunk_mult", 0)
    pack = {"dim": host["dim"], "alpha": host["alpha"], "V": host["V"], "L": host["L"], "w": host["w"], "b": host["b"],
            "v": host["v"], "eig": host["eig"], "order": 1, "mu": host["mu"], "Lq": host["Lq"], "sigma2": host["sigma2"], "lik": lik}
    oval, og = O.elbo_and_grads(pack, args[0].cpu().numpy(), args[1].cpu().numpy(), data_only=True)
    assert abs(out["data_term"] - oval) <= TOL * abs(oval)
    nF = len(host["V"])
    ref = {"mu": og["mu"], "sigma": og["Lq"], "eig": og["eig"], "w": og["w"], "b": og["b"], "v": og["v"],
           "vhat": np.concatenate([og[f"V{i}"] for i in range(nF)], 0),
           "lcat": np.concatenate([np.tril(og[f"L{i}"]).ravel() for i in range(nF)])}
    if lik == "gaussian":
        ref["sigma2"] = og["sigma2"]
    for k, r in ref.items():
        assert rel_err(out[k].reshape(np.shape(r)), r) < TOL, k
        assert elem_err(out[k].reshape(np.shape(r)), r) < 1e-9, k


def test_empty_input_gives_zero_sums():
    prob = _problem(64)
    hp, args = prob["hot"], prob["args"]
    out = hp.elbo_valgrad_packed(args[0][:0], args[1][:0], *args[2:])
    assert float(out.abs().max()) == 0.0


def test_full_size_invariant_at_init_and_determinism():
    """with mu = 0, L_q = I the data term is sum_n[-1/2 log 2 pi s2 - (y^2 + v r^2)/(2 s2)],
    independent of Phi — checked at BASELINE.json's full N = 10 M rows of the headline shape; and two runs are
    bit-identical."""
    N = 10_000_000
    prob = _problem(N)
    hp, args = prob["hot"], list(prob["args"])
    M = hp.M
    args[8] = torch.zeros((M, 1), dtype=torch.float64, device="cuda")
    args[9] = torch.eye(M, dtype=torch.float64, device="cuda")[None]
    out1 = hp.elbo_valgrad_packed(*args).clone()
    out2 = hp.elbo_valgrad_packed(*args).clone()
    assert torch.equal(out1, out2)
    X, Y, w, b, v, s2 = args[0], args[1], args[2], args[3], args[4], args[10]
    r2 = (X * X * w).sum(1) + b
    expect = (-0.5 * math.log(2 * math.pi) - 0.5 * torch.log(s2) - 0.5 * (Y[:, 0] ** 2 + v * r2) / s2).sum()
    assert abs(float(out1[0]) - float(expect)) <= 1e-10 * abs(float(expect))
    # d/dmu of the data term at mu = 0 is A y / s2: compare one strided subsample against the oracle
    idx = torch.arange(0, N, 997, device="cuda")
    sub = hp.elbo_valgrad_packed(X[idx].contiguous(), Y[idx].contiguous(), *args[2:])
    host = prob["host"]
    pack = {"dim": host["dim"], "alpha": host["alpha"], "V": host["V"], "L": host["L"], "w": host["w"], "b": host["b"],
            "v": host["v"], "eig": host["eig"], "order": 1, "mu": np.zeros((M, 1)), "Lq": np.eye(M)[None],
            "sigma2": host["sigma2"], "lik": "gaussian"}
    oval, og = O.elbo_and_grads(pack, X[idx].cpu().numpy(), Y[idx].cpu().numpy(), data_only=True)
    u = hp.unpack(sub)
    assert abs(float(u["data_term"]) - oval) <= TOL * abs(oval)
    assert rel_err(u["mu"].cpu().numpy(), og["mu"]) < TOL
    assert rel_err(u["sigma"].cpu().numpy(), og["Lq"]) < TOL


@pytest.mark.parametrize("T,P,lik,d,F", [(30, 1, "gaussian", 8, 11), (64, 1, "gaussian", 8, 11), (30, 2, "bernoulli", 8, 11),
                                         (12, 5, "gaussian", 8, 11),
                                         (30, 1, "gaussian", 8, 13),            # BASELINE config 4: degrees 0..12 (M = 340)
                                         (64, 1, "bernoulli", 32, 5),           # config 5: d = 32, degrees 0..4, Bernoulli (M = 226)
                                         (2**31 - 1, 1, "gaussian", 2, 11)])    # config 1: S^2, degrees 0..10, no truncation (M = 121)
def test_headline_shapes_match_oracle_on_a_sample(T, P, lik, d, F):
    """The shapes of BASELINE configs 1-5 (d = 8 degrees 0..10 with T = 30 / 64, degrees 0..12, d = 32 Bernoulli, S^2) and
    P = 2 / P = 5 variants: 3001 rows (ragged: not a multiple of any tile) against the oracle, every gradient."""
    N = 3001
    prob = _problem(N, T=T, P=P, lik=lik, d=d, F=F)
    hp, args, host = prob["hot"], prob["args"], prob["host"]
    out = hp.unpack(hp.elbo_valgrad_packed(*args))
    pack = {"dim": host["dim"], "alpha": host["alpha"], "V": host["V"], "L": host["L"], "w": host["w"], "b": host["b"],
            "v": host["v"], "eig": host["eig"], "order": 1, "mu": host["mu"], "Lq": host["Lq"], "sigma2": host["sigma2"], "lik": lik}
    oval, og = O.elbo_and_grads(pack, args[0].cpu().numpy(), args[1].cpu().numpy(), data_only=True)
    assert abs(float(out["data_term"]) - oval) <= TOL * abs(oval)
    F = len(host["V"])
    cmp = {"mu": og["mu"], "sigma": og["Lq"], "eig": og["eig"], "w": og["w"], "b": og["b"], "v": og["v"],
           "vhat": np.concatenate([og[f"V{i}"] for i in range(F)], 0),
           "lcat": np.concatenate([np.tril(og[f"L{i}"]).ravel() for i in range(F)])}
    if lik == "gaussian":
        cmp["sigma2"] = og["sigma2"]
    for k, ref in cmp.items():
        assert rel_err(out[k].cpu().numpy(), ref) < TOL, k
    fmu, fvar = hp.predict_diag(*[args[i] for i in (0, 2, 3, 4, 5, 6, 7, 8, 9)])
    ofmu, ofvar = O.predict_diag(O.pack_to_torch(pack), O.T(args[0].cpu().numpy()))
    assert rel_err(fmu.cpu().numpy(), ofmu.numpy()) < TOL and rel_err(fvar.cpu().numpy(), ofvar.numpy()) < TOL
    # the explicit projection A = sqrt(lambda v) r Phi on the same ragged rows (odd N: unaligned output rows)
    A, r = hp.features(args[0], args[6], args[7], w=args[2], b=args[3], degree_scale=torch.sqrt(args[5] * args[4]), mul_r=True)
    oA = O.projection(O.pack_to_torch(pack), O.T(args[0].cpu().numpy()))
    assert rel_err(A.cpu().numpy(), oA.numpy()) < TOL


@pytest.mark.parametrize("N,T,P,lik,d,F", [(1000, 30, 1, "gaussian", 8, 11), (333, 12, 2, "bernoulli", 5, 5),
                                             (65, 30, 1, "gaussian", 8, 11), (257, 8, 1, "gaussian", 2, 6)])
def test_results_do_not_depend_on_workspace_contents(N, T, P, lik, d, F):
    """compute-sanitizer is closed on this pool, so uninitialised reads are hunted the other way round: the
    caller-provided scratch is poisoned with NaN before every call; every output must stay finite and bit-identical
    to the run on a zeroed workspace (ragged sizes, every entry point that takes scratch)."""
    from gpfy_b200 import _lib
    from gpfy_b200.synthetic import headline_problem
    prob = headline_problem(N, d=d, F=F, T=T, P=P, likelihood=lik, device="cuda")
    hp, args = prob["hot"], prob["args"]
    X, Y, w, b, v, eig, vhat, lcat, mu, sg, s2 = args

    def run_all():
        out = hp.elbo_valgrad_packed(*args).clone()
        fmu, fvar = hp.predict_diag(X, w, b, v, eig, vhat, lcat, mu, sg)
        A, r = hp.features(X, vhat, lcat, w=w, b=b, degree_scale=torch.sqrt(eig * v), mul_r=True)
        pm, pv = hp.project_q(A, mu, sg)
        return [out, fmu.clone(), fvar.clone(), A.clone(), r.clone(), pm.clone(), pv.clone()]

    def poison(val):
        for op in (_lib.OP_FEATURES, _lib.OP_ELBO, _lib.OP_PREDICT):
            ws = hp.workspace(N, op)
            ws.view(torch.float64).fill_(val)

    run_all()               # allocates the scratch buffers
    poison(0.0)
    ref = run_all()
    poison(float("nan"))
    got = run_all()
    for a, b_ in zip(got, ref):
        assert bool(torch.isfinite(a).all())
        assert torch.equal(a, b_)


@pytest.mark.parametrize("name", ["d8_F6_T12", "d8_F11_T30", "d32_F4_T16_bernoulli", "d5_ntk3_P2"])
def test_inducing_basis_on_device_matches_reference_and_autograd(name):
    """gpfy_inducing_basis_fwd / _vjp against the reference's Vs / Ls outputs and against
    torch.autograd through normalise -> Gram -> Cholesky of the oracle."""
    from gpfy_b200.ops import HotPath
    g = load(name)
    F = int(g["meta_F"])
    alpha = float(g["meta_alpha"])
    V = [g[f"p_V_{n:02d}"] for n in range(F)]
    m = [v.shape[0] for v in V]
    hp = HotPath(V[0].shape[1] - 1, m)
    vhat, lcat = hp.inducing_basis(np.concatenate(V, 0), normalise=True)
    assert rel_err(vhat.cpu().numpy(), np.concatenate([g[f"o_Vs_{n:02d}"] for n in range(F)], 0)) < 1e-14
    assert rel_err(lcat.cpu().numpy(), np.concatenate([np.tril(g[f"o_Ls_{n:02d}"]).ravel() for n in range(F)])) < 1e-10
    rng = np.random.default_rng(3)
    Vt = [O.T(v).requires_grad_(True) for v in V]
    Vh = O.normalise_V(Vt)
    Ls = O.orthogonalise_basis(Vh, alpha)
    dvh = [rng.standard_normal(v.shape) for v in V]
    dL = [np.tril(rng.standard_normal((k, k))) for k in m]
    s = sum((v * O.T(d)).sum() for v, d in zip(Vh, dvh)) + sum((l * O.T(d)).sum() for l, d in zip(Ls, dL))
    grads = torch.autograd.grad(s, Vt)
    dV = hp.inducing_basis_vjp(np.concatenate(V, 0), lcat, np.concatenate(dvh, 0), np.concatenate([d.ravel() for d in dL]), normalise=True)
    assert rel_err(dV.cpu().numpy(), np.concatenate([x.numpy() for x in grads], 0)) < 1e-10


def test_inducing_basis_on_device_beyond_64_points_per_degree():
    """m_l = 100 (phase_truncation = 100 of the reference's sweeps, spherical_harmonics.py:119-127): the _ws entry points keep the four
    m_l x m_l matrices of a degree in global scratch instead of shared memory.  Forward and reverse pass against the oracle
    (normalise -> Gram -> Gegenbauer -> Cholesky and torch.autograd through it)."""
    from gpfy_b200.ops import HotPath
    rng = np.random.default_rng(11)
    dim, m, alpha = 9, [1, 9, 30, 100, 100, 70], 3.5   # N(9, 2) = 44 harmonics of degree 2: 30 points keep that Gram matrix regular
    V = [rng.standard_normal((k, dim)) for k in m]
    hp = HotPath(dim - 1, m)
    vhat, lcat = hp.inducing_basis(np.concatenate(V, 0), normalise=True)
    Vt = [O.T(v).requires_grad_(True) for v in V]
    Vh = O.normalise_V(Vt)
    Ls = O.orthogonalise_basis(Vh, alpha)
    assert rel_err(vhat.cpu().numpy(), np.concatenate([v.detach().numpy() for v in Vh], 0)) < 1e-14
    assert rel_err(lcat.cpu().numpy(), np.concatenate([np.tril(l.detach().numpy()).ravel() for l in Ls])) < 1e-10
    dvh = [rng.standard_normal(v.shape) for v in V]
    dL = [np.tril(rng.standard_normal((k, k))) for k in m]
    s = sum((v * O.T(d)).sum() for v, d in zip(Vh, dvh)) + sum((l * O.T(d)).sum() for l, d in zip(Ls, dL))
    grads = torch.autograd.grad(s, Vt)
    dV = hp.inducing_basis_vjp(np.concatenate(V, 0), lcat, np.concatenate(dvh, 0), np.concatenate([d.ravel() for d in dL]), normalise=True)
    assert rel_err(dV.cpu().numpy(), np.concatenate([x.numpy() for x in grads], 0)) < 1e-10


def test_fused_and_general_kernel_paths_agree_on_the_headline_shape():
    """Three independent implementations of the same step: the headline kernels (tensor-core zonal forward + zonal_bwd3,
    P = 1, ambient dim <= 16, Mp <= 288), the general fused backward (zonal_bwd4: C + derivative panel, any Mp / dim) and
    the legacy one-point-per-lane kernels that serve P > 1 and projection weights, selected by the library's development
    switches.  Same inputs -> same packed value and gradients, features and predictions."""
    from gpfy_b200 import ops
    prob = _problem(20_011)
    hp, args = prob["hot"], prob["args"]
    pred_args = [args[i] for i in (0, 2, 3, 4, 5, 6, 7, 8, 9)]
    ds = torch.sqrt(args[5] * args[4])
    fused = hp.elbo_valgrad_packed(*args).clone()
    fmu, fvar = (t.clone() for t in hp.predict_diag(*pred_args))
    A = hp.features(args[0], args[6], args[7], w=args[2], b=args[3], degree_scale=ds, mul_r=True)[0].clone()
    all_opts = ("no_fused_bwd", "no_bwd4", "old_fwd", "old_features")
    lay = hp.layout
    names = ("data_term", "d_mu", "d_sigma", "d_eig", "d_vhat", "d_lcat", "d_w", "d_b", "d_v", "d_sigma2")

    def same(x, y, tol):
        for name in names:
            lo = getattr(lay, name)
            nxt = min([getattr(lay, n) for n in names[1:] + ("total",) if getattr(lay, n) > lo])
            assert rel_err(x[lo:nxt].cpu().numpy(), y[lo:nxt].cpu().numpy()) < tol, name

    try:
        ops.debug_option("no_fused_bwd", 1)                      # -> zonal_bwd4 (+ lik_point + GEMM psi epilogue)
        same(hp.elbo_valgrad_packed(*args), fused, 1e-11)
        for k in all_opts:                                       # -> legacy kernels everywhere
            ops.debug_option(k, 1)
        same(hp.elbo_valgrad_packed(*args), fused, 1e-11)
        gmu, gvar = hp.predict_diag(*pred_args)
        gA = hp.features(args[0], args[6], args[7], w=args[2], b=args[3], degree_scale=ds, mul_r=True)[0]
    finally:
        for k in all_opts:
            ops.debug_option(k, 0)
    assert rel_err(fmu.cpu().numpy(), gmu.cpu().numpy()) < 1e-12 and rel_err(fvar.cpu().numpy(), gvar.cpu().numpy()) < 1e-12
    assert rel_err(A.cpu().numpy(), gA.cpu().numpy()) < 1e-12
