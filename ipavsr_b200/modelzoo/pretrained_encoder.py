"""DBNF encoder stack builder — mirrors `modelzoo/pretrained_encoder.py:4-16` of the reference."""
from ..layers import DenseLayer


def create_pretrained_encoder(incoming, weights, biases, shapes, nonlinearities, names):
    encoder = incoming
    for i, num_units in enumerate(shapes):
        encoder = DenseLayer(encoder, num_units, W=weights[i], b=biases[i],
                             nonlinearity=nonlinearities[i], name=names[i])
    return encoder


def create_encoder(incoming, shapes, nonlinearities, names):
    encoder = incoming
    for i, num_units in enumerate(shapes):
        encoder = DenseLayer(encoder, num_units, nonlinearity=nonlinearities[i], name=names[i])
    return encoder


def extract_dbn_weights(dbn):
    """Legacy encoder form: any object with `get_all_layers()` whose [1..4] carry `.W/.b`
    (`modelzoo/deltanet.py:63-73`, `adenet_v3.py:48-61`)."""
    layers = dbn.get_all_layers()
    weights = [_val(layers[i].W) for i in (1, 2, 3, 4)]
    biases = [_val(layers[i].b) for i in (1, 2, 3, 4)]
    return weights, biases


def _val(x):
    import numpy as np
    if hasattr(x, 'get_value'):
        x = x.get_value()
    return np.asarray(x).astype('float32')
