"""Checkpoint / weight-file IO — mirrors `utils/io.py:18-47` and the loaders of the runners, without Lasagne.

Formats kept byte-compatible:
  (i)   pickle of `list[np.ndarray]` from `get_all_param_values` (`utils/io.py:40-47`; Python-2 pickles load with
        encoding='latin1');
  (ii)  encoder `.mat`: keys `w1..wN` (in,out), `b1..bN` (1,out) (`runners/2stream_dct.py:31-40`);
  (iii) LSTM `.mat`: `{prefix}_w_{in,hid}_to_{ingate,forgetgate,cell,outgate}`, `{prefix}_b_{...}`
        (`modelzoo/deltanet_majority_vote.py:158-196`, consumed by `custom/layers.py:28-52`).
"""
import pickle

import numpy as np
import scipy.io as sio

from .. import layers as L
from ..custom.nonlinearities import select_nonlinearity


def read_data_split_file(path, sep=','):
    with open(path) as f:
        return [int(s) for s in f.readline().split(sep)]


def load_mat_file(path):
    return sio.loadmat(path)


def save_mat(d, path):
    sio.savemat(path, d)


def save_model(model, path):
    with open(path, 'wb') as f:
        pickle.dump(model, f)


def load_model(path):
    with open(path, 'rb') as f:
        try:
            return pickle.load(f)
        except UnicodeDecodeError:
            f.seek(0)
            return pickle.load(f, encoding='latin1')


def save_model_params(network, path):
    """`utils/io.py:40-42`: pickle of get_all_param_values(network) (protocol 2 so Python 2 can read it back)."""
    with open(path, 'wb') as f:
        pickle.dump(L.get_all_param_values(network), f, protocol=2)


def load_model_params(network, path):
    """`utils/io.py:45-47`."""
    L.set_all_param_values(network, load_model(path))
    return network


def load_decoder(path, shapes, nonlinearities):
    """`runners/2stream_dct.py:31-40`: encoder 4-tuple (weights, biases, shapes, nonlinearities) from a `.mat`."""
    nn = sio.loadmat(path) if isinstance(path, str) else path
    shapes = [int(s) for s in shapes.split(',')] if isinstance(shapes, str) else list(shapes)
    if isinstance(nonlinearities, str):
        nonlinearities = [select_nonlinearity(n.strip()) for n in nonlinearities.split(',')]
    weights, biases = [], []
    for i in range(len(shapes)):
        weights.append(np.asarray(nn['w{}'.format(i + 1)]).astype('float32'))
        biases.append(np.asarray(nn['b{}'.format(i + 1)])[0].astype('float32'))
    return weights, biases, shapes, nonlinearities


def save_decoder(path, weights, biases):
    d = {}
    for i, (w, b) in enumerate(zip(weights, biases)):
        d['w{}'.format(i + 1)] = np.asarray(w, 'float32')
        d['b{}'.format(i + 1)] = np.asarray(b, 'float32').reshape(1, -1)
    sio.savemat(path, d)
