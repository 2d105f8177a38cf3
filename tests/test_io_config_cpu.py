"""CPU tests: checkpoint formats (pickle list, encoder .mat, LSTM .mat) and .ini option parsing stay drop-in."""
import io
import os
import pickle

import numpy as np
import pytest

from ipavsr_b200 import layers as L, config, nonlinearities as nl
from ipavsr_b200.utils import io as uio
from ipavsr_b200.modelzoo import deltanet_majority_vote as dmv
from ipavsr_b200.custom.layers import create_pretrained_lstm
from ipavsr_b200.layers import Gate, InputLayer
import model_util as MU


def test_param_pickle_roundtrip(tmp_path):
    rng = np.random.default_rng(0)
    s = MU.build('adenet_v2', rng, fusiontype='adasum')
    MU.randomize_params(s['net'], rng)
    vals = L.get_all_param_values(s['net'])
    path = str(tmp_path / 'best.pkl')
    uio.save_model_params(s['net'], path)
    loaded = pickle.load(open(path, 'rb'))
    assert isinstance(loaded, list) and len(loaded) == len(vals)
    assert all(isinstance(a, np.ndarray) and a.dtype == np.float32 for a in loaded)
    s2 = MU.build('adenet_v2', np.random.default_rng(1), fusiontype='adasum')
    uio.load_model_params(s2['net'], path)
    for a, b in zip(vals, L.get_all_param_values(s2['net'])):
        np.testing.assert_array_equal(a, b)
    # a Python-2 style (protocol 2, latin1) pickle loads too
    with open(path, 'wb') as f:
        pickle.dump([v for v in vals], f, protocol=2)
    assert len(uio.load_model(path)) == len(vals)


def test_encoder_mat_roundtrip(tmp_path):
    rng = np.random.default_rng(2)
    W, b = MU.enc_weights(rng, 40, (60, 40, 30, 10))
    path = str(tmp_path / 'enc.mat')
    uio.save_decoder(path, W, b)
    w2, b2, shapes, nonlins = uio.load_decoder(path, '60,40,30,10', 'rectify,rectify,rectify,linear')
    assert shapes == [60, 40, 30, 10] and [n.name for n in nonlins] == ['rectify'] * 3 + ['linear']
    for a, c in zip(W + b, w2 + b2):
        np.testing.assert_array_equal(a, c)
    assert b2[0].shape == (60,)


def test_lstm_mat_extract_and_reload():
    rng = np.random.default_rng(3)
    s = MU.build('deltanet_majority_vote', rng)
    MU.randomize_params(s['net'], rng)
    d = dmv.extract_lstm_weights(s['net'], ['f_blstm1', 'b_blstm1'], ['f', 'b'])
    assert len(d) == 24 and d['f_w_in_to_cell'].shape == (30, 12) and d['b_b_outgate'].shape == (12,)
    enc = dmv.extract_encoder_weights(s['net'], ['fc1', 'bottleneck'], [('w1', 'b1'), ('w4', 'b4')])
    assert enc['w1'].shape == (40, 60) and enc['b4'].shape == (10,)
    g = Gate()
    l_in, l_mask = InputLayer((None, None, 30)), InputLayer((None, None))
    lstm = create_pretrained_lstm(d, 'f', l_in, l_mask, 12, Gate(W_cell=None, nonlinearity=nl.tanh), g, 'lstm_x')
    f = [l for l in L.get_all_layers(s['net']) if l.name == 'f_blstm1'][0]
    np.testing.assert_array_equal(lstm.W_hid_to_forgetgate.get_value(), f.W_hid_to_forgetgate.get_value())
    np.testing.assert_array_equal(lstm.b_cell.get_value(), f.b_cell.get_value())


INI_MODERN = u"""
[stream1]
data = a.mat
model = m.mat
shape = 2000,1000,500,50
nonlinearities = rectify,rectify,rectify,linear
input_dimensions = 1200
samplewisenormalize = True
[stream2]
data = b.mat
input_dimensions = 90
has_encoder = False
[lstm_classifier]
fusiontype = concat
weight_init = ortho
use_peepholes = True
windowsize = 9
output_classes = 10
lstm_size = 250
[training]
learning_rate = 0.0001
num_epoch = 40
batchsize = 10
"""
INI_LEGACY = u"""
[models]
fusiontype = adasum
[training]
learning_rate = 2.0
decay_rate = 0.8
decay_start = 20
lstm_units = 250
output_units = 26
do_finetune = False
"""


def test_ini_options_both_generations():
    m = config.model_options(config.read(io.StringIO(INI_MODERN)))
    assert (m['fusiontype'], m['lstm_size'], m['output_classes'], m['use_peepholes'], m['windowsize']) == \
        ('concat', 250, 10, True, 9)
    assert len(m['streams']) == 2 and m['streams'][0]['input_dimensions'] == 1200 and not m['streams'][1]['has_encoder']
    assert type(config.weight_init_fn(m['weight_init'])).__name__ == 'Orthogonal'
    t = config.training_options(config.read(io.StringIO(INI_MODERN)))
    assert t['learning_rate'] == 1e-4 and t['batchsize'] == 10
    m2 = config.model_options(config.read(io.StringIO(INI_LEGACY)))
    assert (m2['fusiontype'], m2['lstm_size'], m2['output_classes']) == ('adasum', 250, 26)
    t2 = config.training_options(config.read(io.StringIO(INI_LEGACY)))
    assert t2['learning_rate'] == 2.0 and t2['decay_rate'] == 0.8 and t2['decay_start'] == 20 and not t2['do_finetune']
    with pytest.raises(ValueError):
        config.weight_init_fn('bogus')
