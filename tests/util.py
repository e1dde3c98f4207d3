import numpy as np

P = 2013265921


def rand_field(rng, shape):
    return rng.integers(0, P, size=shape, dtype=np.uint32)


def bitrev(x, bits):
    r = 0
    for i in range(bits):
        r = (r << 1) | ((x >> i) & 1)
    return r


def bitrev_perm(n_bits):
    return np.array([bitrev(i, n_bits) for i in range(1 << n_bits)], dtype=np.int64)
