"""Synthetic stand-in for the reference's `datasets` package (MAF-style UCI loaders that read files this image does
not have): the same attribute layout (`.trn.x`, `.val.x`, `.tst.x` float32 arrays) filled with seeded correlated
Gaussians of the right dimensionality."""
import numpy as np


class _Split:
    def __init__(self, x):
        self.x = x
        self.N = x.shape[0]


class _Synthetic:
    def __init__(self, dim, seed, n_trn=300, n_val=100, n_tst=100):
        rng = np.random.RandomState(seed)
        mix = rng.standard_normal((dim, dim)).astype(np.float32) / np.sqrt(dim)

        def draw(n):
            z = rng.standard_normal((n, dim)).astype(np.float32)
            return (z @ mix + 0.1 * z ** 2).astype(np.float32)
        self.trn, self.val, self.tst = _Split(draw(n_trn)), _Split(draw(n_val)), _Split(draw(n_tst))
        self.n_dims = dim


def POWER():
    return _Synthetic(6, 0)


def GAS():
    return _Synthetic(8, 1)


def HEPMASS():
    return _Synthetic(21, 2)


def MINIBOONE():
    return _Synthetic(43, 3)


def BSDS300():
    return _Synthetic(63, 4, n_trn=64, n_val=32, n_tst=32)
