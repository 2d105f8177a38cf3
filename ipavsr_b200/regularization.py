"""lasagne.regularization, the subset nolearn's `objective_l2` uses (`avletters/trimodal.py:87`):
`loss + objective_l2 * regularize_network_params(output_layer, l2)` — the sum of squares of every regularizable parameter
(weights; not biases, initial states or BatchNorm statistics).  Evaluated on the device by `ipavsr_l2_penalty`."""


class Penalty(object):
    def __init__(self, layer, kind='l2', coefficient=1.0):
        self.layer, self.kind, self.coefficient = layer, kind, float(coefficient)

    def __mul__(self, c):
        return Penalty(self.layer, self.kind, self.coefficient * float(c))

    __rmul__ = __mul__


def l2(x):
    raise TypeError('l2 is only usable as the penalty argument of regularize_network_params / regularize_layer_params')


def regularize_network_params(layer, penalty, tags=None, **kwargs):
    if penalty is not l2:
        raise NotImplementedError('only the l2 penalty is implemented')
    if tags not in (None, {'regularizable': True}):
        raise NotImplementedError('only the regularizable tag is supported')
    return Penalty(layer)


def regularize_layer_params(layer, penalty, tags=None, **kwargs):
    """nolearn passes every layer of the net; the last one reaches all of them."""
    if isinstance(layer, (list, tuple)):
        layer = layer[-1]
    return regularize_network_params(layer, penalty, tags, **kwargs)
