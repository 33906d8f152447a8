"""Independent numpy restatement of the epoch permutation of gpfy_minibatch_gather (include/gpfy_b200.h): splitmix64
round keys, a 6-round balanced Feistel network on 2 * half_bits bits with a murmur3-finaliser round function, cycle-walked
into [0, n).  Test infrastructure only."""
import numpy as np

ROUNDS = 6
M64 = (1 << 64) - 1


def round_keys(seed, epoch):
    z = (seed * 0x9E3779B97F4A7C15 + epoch * 0xD1B54A32D192ED03 + 0x2545F4914F6CDD1D) & M64
    keys = []
    for _ in range(ROUNDS):
        z = (z + 0x9E3779B97F4A7C15) & M64
        y = z
        y = ((y ^ (y >> 30)) * 0xBF58476D1CE4E5B9) & M64
        y = ((y ^ (y >> 27)) * 0x94D049BB133111EB) & M64
        y ^= y >> 31
        keys.append(y >> 32)
    return keys


def half_bits(n):
    bits = 2
    while bits < 62 and (1 << bits) < n:
        bits += 1
    return (bits + 1) // 2


def _round(r, key):
    h = (r.astype(np.uint64) * np.uint64(0x9E3779B1) + np.uint64(key)) & np.uint64(0xFFFFFFFF)
    h ^= h >> np.uint64(16)
    h = (h * np.uint64(0x85EBCA6B)) & np.uint64(0xFFFFFFFF)
    h ^= h >> np.uint64(13)
    h = (h * np.uint64(0xC2B2AE35)) & np.uint64(0xFFFFFFFF)
    h ^= h >> np.uint64(16)
    return h


def permutation(seed, epoch, n, positions=None):
    """pi(positions) (default: all of [0, n)) as int64."""
    hb = half_bits(n)
    mask = np.uint64((1 << hb) - 1)
    keys = round_keys(seed, epoch)
    v = (np.arange(n) if positions is None else np.asarray(positions)).astype(np.uint64)
    todo = np.ones(v.shape, dtype=bool)
    while todo.any():
        x = v[todo]
        l, r = x >> np.uint64(hb), x & mask
        for k in keys:
            l, r = r, l ^ (_round(r & np.uint64(0xFFFFFFFF), k) & mask)
        x = (l << np.uint64(hb)) | r
        v[todo] = x
        todo[todo] = x >= np.uint64(n)
    return v.astype(np.int64)
