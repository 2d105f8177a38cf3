"""Weight initialisers with Lasagne's names and distributions.

The reference passes `las.init.Orthogonal()`, `GlorotUniform()`, `Normal(0.1)`, `Uniform()` and
`Constant(0.)` objects into `Gate(...)` (`runners/2stream_dct.py:187-195`, `modelzoo/adenet_v2.py:20-28`).
RNG is the global `numpy.random`, as in Lasagne, so seeding `numpy.random.seed` makes builds reproducible.
"""
import numpy as np


class Initializer(object):
    def __call__(self, shape):
        return self.sample(tuple(int(s) for s in shape))


class Constant(Initializer):
    def __init__(self, val=0.0):
        self.val = val

    def sample(self, shape):
        return np.full(shape, self.val, dtype=np.float32)


class Normal(Initializer):
    def __init__(self, std=0.01, mean=0.0):
        self.std, self.mean = std, mean

    def sample(self, shape):
        return np.random.normal(self.mean, self.std, size=shape).astype(np.float32)


class Uniform(Initializer):
    def __init__(self, range=0.01, std=None, mean=0.0):
        if std is not None:
            a = mean - np.sqrt(3) * std
            b = mean + np.sqrt(3) * std
        else:
            try:
                a, b = range
            except TypeError:
                a, b = -range, range
        self.range = (a, b)

    def sample(self, shape):
        return np.random.uniform(self.range[0], self.range[1], size=shape).astype(np.float32)


class GlorotUniform(Initializer):
    def __init__(self, gain=1.0):
        self.gain = np.sqrt(2) if gain == 'relu' else gain

    def sample(self, shape):
        if len(shape) < 2:
            raise RuntimeError('GlorotUniform only works with shapes of length >= 2')
        n1, n2 = shape[:2]
        rf = int(np.prod(shape[2:]))
        std = self.gain * np.sqrt(2.0 / ((n1 + n2) * rf))
        a = np.sqrt(3) * std
        return np.random.uniform(-a, a, size=shape).astype(np.float32)


class Orthogonal(Initializer):
    def __init__(self, gain=1.0):
        self.gain = np.sqrt(2) if gain == 'relu' else gain

    def sample(self, shape):
        if len(shape) < 2:
            raise RuntimeError('Only shapes of length 2 or more are supported.')
        flat = (shape[0], int(np.prod(shape[1:])))
        a = np.random.normal(0.0, 1.0, flat)
        u, _, v = np.linalg.svd(a, full_matrices=False)
        q = u if u.shape == flat else v
        return (self.gain * q.reshape(shape)).astype(np.float32)
