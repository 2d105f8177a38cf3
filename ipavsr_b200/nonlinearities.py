"""Nonlinearity tokens for the .ini `nonlinearities:` option.

Mirrors the string table of the reference (`custom/nonlinearities.py:4-16`, which re-exports
`lasagne.nonlinearities`).  A nonlinearity here is an opaque token carrying the integer code the CUDA
epilogue switches on (see `include/ipavsr_b200.h`, `IPAVSR_ACT_*`); there is no Python arithmetic behind it.
"""


class Nonlinearity(object):
    def __init__(self, name, code):
        self.name = name
        self.code = code

    def __repr__(self):
        return 'nonlinearity<%s>' % self.name


linear = Nonlinearity('linear', 0)
identity = linear
sigmoid = Nonlinearity('sigmoid', 1)
rectify = Nonlinearity('rectify', 2)
tanh = Nonlinearity('tanh', 3)
leaky_rectify = Nonlinearity('leaky_rectify', 4)            # slope 0.01
very_leaky_rectify = Nonlinearity('very_leaky_rectify', 5)  # slope 1/3
softplus = Nonlinearity('softplus', 6)
elu = Nonlinearity('elu', 7)
softmax = Nonlinearity('softmax', 8)                        # row-wise; only valid on the head Dense


def select_nonlinearity(string):
    """String -> nonlinearity (reference `custom/nonlinearities.py:4`).  Unknown keys raise KeyError
    exactly like the reference's dict lookup; `scaled_tanh` (a class in Lasagne, never used by a shipped
    config) is not supported and raises ValueError."""
    table = {'rectify': rectify, 'sigmoid': sigmoid, 'leaky_rectify': leaky_rectify,
             'very_leaky_rectify': very_leaky_rectify, 'tanh': tanh, 'linear': linear,
             'softmax': softmax, 'softplus': softplus, 'elu': elu, 'identity': identity}
    if string == 'scaled_tanh':
        raise ValueError('scaled_tanh is not supported by the B200 path')
    return table[string]
