from .. import init
from ..layers import Gate
from ..nonlinearities import tanh


def gates(w_init_fn):
    """The (gate_parameters, cell_parameters) pair every builder creates (`modelzoo/adenet_v2.py:20-28`)."""
    gate_parameters = Gate(W_in=w_init_fn, W_hid=w_init_fn, b=init.Constant(0.))
    cell_parameters = Gate(W_in=w_init_fn, W_hid=w_init_fn, W_cell=None, b=init.Constant(0.), nonlinearity=tanh)
    return gate_parameters, cell_parameters


def fuse(fusiontype, incomings, names, strict):
    """sum | adasum | concat fusion.  `strict` builders raise ValueError on an unknown type (the reference's
    `raise ValueError(message=...)`, `adenet_v2.py:75`); the others leave `l_fuse` undefined in the reference
    (a NameError, `adenet_v3.py:147-154`) — here every builder raises ValueError."""
    from ..layers import ElemwiseSumLayer, ConcatLayer, AdaptiveElemwiseSumLayer
    if fusiontype == 'sum':
        return ElemwiseSumLayer(incomings, name=names['sum'])
    if fusiontype == 'adasum':
        return AdaptiveElemwiseSumLayer(incomings, name=names['adasum'])
    if fusiontype == 'concat':
        return ConcatLayer(incomings, axis=-1, name=names['concat'])
    raise ValueError('Unsupported Fusion Type used!')
