"""The arithmetic of the int8 digit-product contraction (csrc/gemm_i8.cuh), restated in numpy: the signed radix-256 split, its exactness
and the truncation of the product at digit groups i + j <= 7.  Pins the error model the kernel's header states (and DESIGN.md §4c)
without a GPU: normwise error of C = W2 Z well below 1e-12 for operands with the dynamic range of the real ones."""
import numpy as np

S = 6


def digits(x):
    """x (|x| <= 0.4902) -> S signed digits d_i in [-128, 127] with x = sum_i d_i 256^-i + e, |e| <= 2^-49 (gi8_digits: one FMA with
    1.5 * 2^52, the unsigned-digit trick T = q + 0x80..80)."""
    q = np.rint(np.asarray(x, dtype=np.float64) * 2.0 ** 48).astype(np.int64)
    t = (q + 0x808080808080).astype(np.uint64) ^ np.uint64(0x808080808080)
    return [((t >> np.uint64(8 * (S - 1 - i))) & np.uint64(0xFF)).astype(np.uint8).view(np.int8).astype(np.int64) for i in range(S)]


def pow2_at_least(v):
    return 2.0 ** np.ceil(np.log2(v))


def test_digit_split_is_exact_to_2_pow_minus_49():
    rng = np.random.default_rng(0)
    x = np.concatenate([rng.uniform(-0.4902, 0.4902, 100_000), [0.0, 0.4902, -0.4902, 2.0 ** -48, -(2.0 ** -49), 127.0 / 256 + 1e-9]])
    d = digits(x)
    assert all(di.min() >= -128 and di.max() <= 127 for di in d)
    rec = sum(di.astype(np.float64) * 256.0 ** -(i + 1) for i, di in enumerate(d))
    assert np.max(np.abs(rec - x)) <= 2.0 ** -49


def contraction_i8(W, Z, bound):
    """C = W Z the way the kernel forms it: per-feature scales s_k >= 2.04 bound_k (powers of two), per-row scales r_o, digit planes,
    exact integer products of the 21 digit pairs with i + j <= 7, float64 recombination."""
    s = pow2_at_least(2.04 * bound)
    Ws = W * s[None, :]
    r = pow2_at_least(2.04 * np.maximum(np.abs(Ws).max(1), 1e-300))
    a = digits(Z / s[:, None])              # [k, n] digits of z'
    b = digits(Ws / r[:, None])             # [o, k] digits of w'
    acc = [np.zeros((W.shape[0], Z.shape[1]), dtype=np.int64) for _ in range(S)]
    for i in range(S):
        for j in range(S - i):
            acc[i + j] += b[j] @ a[i]       # exact: |sum| < 2^27
    assert max(int(np.abs(g).max()) for g in acc) < 2 ** 31
    t = np.zeros_like(acc[0], dtype=np.float64)
    for g in range(S - 1, -1, -1):
        t = t / 256.0 + acc[g]
    return t * (r[:, None] / 65536.0)


def test_contraction_error_model():
    rng = np.random.default_rng(1)
    M, n = 288, 512
    # rows of Z bounded like Gegenbauer values of degrees 0..10 (1 .. 8008), W2 with rows spanning six orders of magnitude
    bound = np.repeat([1.0, 7.0, 31.5, 115.5, 375.4, 1126.1, 3003.0, 8008.0, 8008.0], 32)
    Z = rng.uniform(-1, 1, (M, n)) * bound[:, None]
    W = rng.standard_normal((M, M)) / bound[None, :] * 10.0 ** rng.uniform(-3, 3, (M, 1))
    C = contraction_i8(W, Z, bound)
    ref = (W.astype(np.longdouble) @ Z.astype(np.longdouble)).astype(np.float64)
    err = np.abs(C - ref).max(1) / np.abs(ref).max(1)          # normwise per output row
    assert err.max() < 1e-13, err.max()
    # W2 = 0 (the prior, S = I): every digit is zero and C is exactly zero
    assert not contraction_i8(np.zeros((M, M)), Z, bound).any()
