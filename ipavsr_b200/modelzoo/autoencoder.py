"""Auto-encoder builders — mirrors `modelzoo/autoencoder.py:11-59` of the reference (the 8-layer DBNF encoder/decoder
whose first four layers become the stream encoders)."""
import scipy.io as sio

from ..layers import DenseLayer


def load_dbn(path='models/avletters_ae.mat'):
    """`modelzoo/autoencoder.py:11-38`: (weights w1..w8, biases b1..b8) of a MATLAB-pretrained DBN; `path` may also be
    an already loaded dict."""
    nn = sio.loadmat(path) if isinstance(path, str) else path
    n = 0
    while 'w%d' % (n + 1) in nn:
        n += 1
    return [nn['w%d' % (i + 1)] for i in range(n)], [nn['b%d' % (i + 1)][0] for i in range(n)]


def create_model(incoming, weights, biases, activations, layersizes):
    """`modelzoo/autoencoder.py:41-53`: one DenseLayer per weight matrix, named fc1..fcN."""
    for i, w in enumerate(weights):
        incoming = DenseLayer(incoming, layersizes[i], w, biases[i], activations[i], name='fc{}'.format(i + 1))
    return incoming


def create_pretrained_encoder(incoming, weights, biases, activations, layersizes):
    """`modelzoo/autoencoder.py:56-61`."""
    names = ('fc1', 'fc2', 'fc3', 'bottleneck')
    for i in range(4):
        incoming = DenseLayer(incoming, layersizes[i], W=weights[i], b=biases[i], nonlinearity=activations[i],
                              name=names[i])
    return incoming
