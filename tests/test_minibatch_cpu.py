"""Host side of the shuffled-minibatch path: the epoch permutation evaluated by the library
(`gpfy_shuffle_index`, the function the gather kernel runs per row) against an independent numpy restatement, and the
properties a shuffle must have.  Bit-exact (integer work)."""
import numpy as np
import pytest

from gpfy_b200.minibatch import shuffle_index
from tests import shuffle_ref


@pytest.mark.parametrize("n", [1, 2, 3, 4, 5, 17, 64, 1000, 4097])
@pytest.mark.parametrize("seed,epoch", [(0, 0), (7, 3)])
def test_library_permutation_matches_restatement_and_is_a_bijection(n, seed, epoch):
    ref = shuffle_ref.permutation(seed, epoch, n)
    lib = np.array([shuffle_index(seed, epoch, n, i) for i in range(n)])
    assert np.array_equal(lib, ref)
    assert np.array_equal(np.sort(ref), np.arange(n))


def test_large_domain_spot_checks():
    n = (1 << 33) + 12345      # half_bits = 17: rows beyond 2^32
    pos = [0, 1, 2, n // 2, n - 1]
    ref = shuffle_ref.permutation(11, 2, n, pos)
    lib = [shuffle_index(11, 2, n, p) for p in pos]
    assert list(ref) == lib and all(0 <= v < n for v in lib) and len(set(lib)) == len(lib)


def test_epochs_and_seeds_give_different_orders_and_mix_well():
    n = 4096
    a, b, c = (shuffle_ref.permutation(s, e, n) for s, e in [(0, 0), (0, 1), (1, 0)])
    assert (a != b).mean() > 0.99 and (a != c).mean() > 0.99
    # no visible structure: rank correlation with the identity and between epochs is at the 1/sqrt(n) noise level
    ident = np.arange(n)
    for p in (a, b, c):
        assert abs(np.corrcoef(ident, p)[0, 1]) < 5 / np.sqrt(n)
    assert abs(np.corrcoef(a, b)[0, 1]) < 5 / np.sqrt(n)
    # consecutive positions do not map to neighbouring rows
    assert np.mean(np.abs(np.diff(a)) <= 2) < 0.01


def test_bad_arguments_are_reported():
    from gpfy_b200 import _lib
    with pytest.raises(_lib.GpfyError):
        shuffle_index(0, 0, 10, 10)
    with pytest.raises(_lib.GpfyError):
        shuffle_index(0, 0, 0, 0)


def test_streamed_slab_bounds_ramp_and_cover():
    """HotPath.slab_bounds: slabs cover [0, n) exactly, none exceeds the workspace capacity, every slab but the last is a
    whole number of tile rounds, the first one is two rounds (its copy cannot overlap any compute) and no slab is more than
    1.5 x (+ one round) its predecessor, so a slab's copy fits under the processing of the one before."""
    from gpfy_b200.ops import HotPath
    unit = 192 * 148
    for n, cap in ((10_000_000, 909_312), (1_250_000, 625_152), (1, 128), (31, 128), (1000, 1024), (150_001, 150_016),
                   (2_000_000, 681_984), (909_312, 909_312), (909_313, 909_312)):
        b = HotPath.slab_bounds(n, cap, unit)
        sizes = [y - x for x, y in zip(b[:-1], b[1:])]
        assert b[0] == 0 and b[-1] == n and all(0 < s <= max(cap, unit) for s in sizes), (n, cap, b)
        assert all(s % unit == 0 or s == cap for s in sizes[:-1])
        assert sizes[0] <= 2 * unit
        assert all(y <= 1.5 * x + unit for x, y in zip(sizes[:-1], sizes[1:-1] or sizes[1:]))
