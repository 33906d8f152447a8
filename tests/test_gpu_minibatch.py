"""Shuffled minibatches on the device: the gather kernel against numpy indexing with the
independently restated permutation (bit-exact), epoch semantics of the reference's generator (optimization.py:31-58:
disjoint batches, last partial batch dropped, reshuffle), pinned-host residency, and minibatch training on the device
against the host loop."""
import numpy as np
import pytest
import torch

from gpfy_b200 import optimization as opt
from gpfy_b200.minibatch import DeviceMinibatches
from gpfy_b200.training import adam
from tests import shuffle_ref
from tests.api_util import objects_from_golden
from tests.golden_util import load, rel_err

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n,dx,dy,batch", [(1, 3, 1, 1), (37, 8, 1, 5), (1000, 8, 2, 128), (5000, 32, 1, 4999), (4096, 1, 3, 4096)])
@pytest.mark.parametrize("pinned", [False, True])
def test_gather_is_bit_exact_and_epochs_follow_the_reference_semantics(n, dx, dy, batch, pinned):
    rng = np.random.default_rng(n + dx)
    X, Y = rng.standard_normal((n, dx)), rng.standard_normal((n, dy))
    if pinned:   # data set left in pinned host memory, rows read through unified addressing
        Xs, Ys = torch.as_tensor(X).pin_memory(), torch.as_tensor(Y).pin_memory()
    else:
        Xs, Ys = torch.as_tensor(X, device="cuda"), torch.as_tensor(Y, device="cuda")
    mb = DeviceMinibatches(Xs, Ys, batch, seed=5, return_index=True)
    per_epoch = n // batch
    for epoch in range(2):
        perm = shuffle_ref.permutation(5, epoch, n)
        seen = []
        for b in range(per_epoch):
            Xb, Yb = mb.next()
            idx = mb.idx.cpu().numpy()
            assert np.array_equal(idx, perm[b * batch:(b + 1) * batch])
            assert np.array_equal(Xb.cpu().numpy(), X[idx]) and np.array_equal(Yb.cpu().numpy(), Y[idx])
            seen.append(idx)
        seen = np.concatenate(seen)
        assert len(np.unique(seen)) == per_epoch * batch          # batches of an epoch are disjoint
        assert mb.epoch == epoch                                   # ... and the partial batch is dropped at the next call


def test_bad_batch_size_raises():
    from gpfy_b200 import _lib
    X = torch.zeros((10, 2), dtype=torch.float64, device="cuda")
    with pytest.raises(_lib.GpfyError):
        DeviceMinibatches(X, X[:, :1], 11)
    with pytest.raises(_lib.GpfyError):
        DeviceMinibatches(X, X[:, :1], 0)


@pytest.mark.parametrize("name", ["d8_F6_T12", "d32_F4_T16_bernoulli"])
def test_device_minibatch_training_matches_host_loop(name):
    """`DeviceTrainer.fit_minibatch` (gather + step, nothing leaves the device) == `GP.fit` over
    `create_training_step(batch_size=...)` with the same seed, loss for loss, including the dataset_size / batch rescale."""
    from gpfy_b200.device_step import DeviceTrainer
    from gpfy_b200.param import tree_leaves
    g = load(name)
    o = objects_from_golden(g)
    q, lik, m, param = o["q"], o["lik"], o["m"], o["param"]
    X, Y = g["X"], g["Y"]
    K, lr, batch = 9, 3e-2, 16          # N = 64 or 48: crosses an epoch boundary
    step = opt.create_training_step(m, {"x": X, "y": Y}, ("x", "y"), q, lik, batch_size=batch, seed=3)
    host_param, _, host_losses = m.fit(param, step, adam(lr), K, progress_bar=False)
    tr = DeviceTrainer(param, m, q, lik, learning_rate=lr).fit_minibatch(X, Y, batch, K, seed=3)
    assert rel_err(tr.losses(), np.asarray(host_losses, dtype=np.float64)) < 1e-9
    for a, b in zip(tree_leaves(tr.param().params), tree_leaves(host_param.params)):
        assert rel_err(a, b) < 1e-8
    # the rescale: a "minibatch" of all rows in one batch reproduces the full-batch loss of the first step
    full = opt.create_training_step(m, {"x": X, "y": Y}, ("x", "y"), q, lik)
    whole = opt.create_training_step(m, {"x": X, "y": Y}, ("x", "y"), q, lik, batch_size=X.shape[0], seed=1)
    _, _, l_full = m.fit(param, full, adam(lr), 1, progress_bar=False)
    _, _, l_whole = m.fit(param, whole, adam(lr), 1, progress_bar=False)
    assert abs(l_full[0] - l_whole[0]) <= 1e-10 * abs(l_full[0])


@pytest.mark.parametrize("name", ["d8_F6_T12", "s2_deg10_polydecay", "d32_F4_T16_bernoulli"])
def test_cuda_graph_replay_matches_eager_steps(name):
    """One captured step (gather, constrain, basis, ELBO + gradient, KL, chain rules, Adam, advance) replayed from a
    CUDA graph == the eager sequence of C-ABI calls: the step count, epoch and position live on the device."""
    from gpfy_b200.device_step import DeviceTrainer
    from gpfy_b200.param import tree_leaves
    g = load(name)
    o = objects_from_golden(g)
    q, lik, m, param = o["q"], o["lik"], o["m"], o["param"]
    X, Y = g["X"], g["Y"]
    K, lr, batch = 11, 3e-2, 16
    eager = DeviceTrainer(param, m, q, lik, learning_rate=lr).fit_minibatch(X, Y, batch, K, seed=3)
    graph = DeviceTrainer(param, m, q, lik, learning_rate=lr).fit_minibatch(X, Y, batch, K, seed=3, use_graph=True)
    assert graph.count == K and len(graph.losses()) == K
    assert rel_err(graph.losses(), eager.losses()) < 1e-12
    for a, b in zip(tree_leaves(graph.param().params), tree_leaves(eager.param().params)):
        assert rel_err(a, b) < 1e-11
    # full batch, and continuing an existing run (the count carries over into Adam's bias correction)
    Xd, Yd = torch.as_tensor(X, device="cuda"), torch.as_tensor(Y, device="cuda")
    e2 = DeviceTrainer(param, m, q, lik, learning_rate=lr).fit(Xd, Yd, 7, use_graph=False)
    g2 = DeviceTrainer(param, m, q, lik, learning_rate=lr).fit(Xd, Yd, 3, use_graph=True).fit(Xd, Yd, 4, use_graph=True)
    assert g2.count == 7 and rel_err(g2.losses(), e2.losses()) < 1e-12
    # one and two steps (no replay at all / capture without replay)
    g1 = DeviceTrainer(param, m, q, lik, learning_rate=lr).fit(Xd, Yd, 1, use_graph=True).fit(Xd, Yd, 2, use_graph=True)
    assert rel_err(g1.losses(), e2.losses()[:3]) < 1e-12
