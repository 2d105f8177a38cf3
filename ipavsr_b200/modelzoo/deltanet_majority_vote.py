"""DeltaNet, frame-level head (majority vote at evaluation) — mirrors
`modelzoo/deltanet_majority_vote.py:14-66` plus the weight extraction helpers (:137-196)."""
from .. import init
from ..layers import (InputLayer, DenseLayer, ReshapeLayer, ElemwiseSumLayer, DeltaLayer, get_all_layers)
from ..nonlinearities import softmax
from ..custom.layers import create_blstm, create_lstm
from .pretrained_encoder import create_pretrained_encoder
from ._common import gates


def create_model(dbn, input_shape, input_var, mask_shape, mask_var, lstm_size=250, win=None,
                 output_classes=26, w_init_fn=init.GlorotUniform(), use_peepholes=False, use_blstm=True):
    weights, biases, shapes, nonlinearities = dbn
    names = ['fc1', 'fc2', 'fc3', 'bottleneck']
    gate_parameters, cell_parameters = gates(w_init_fn)
    l_in = InputLayer(input_shape, input_var, 'input')
    l_mask = InputLayer(mask_shape, mask_var, 'mask')
    l_reshape1 = ReshapeLayer(l_in, (-1, input_shape[-1]), name='reshape1')
    l_encoder = create_pretrained_encoder(l_reshape1, weights, biases, shapes, nonlinearities, names)
    encoder_len = l_encoder.output_shape[-1]
    l_reshape2 = ReshapeLayer(l_encoder, (None, None, encoder_len), name='reshape2')
    l_delta = DeltaLayer(l_reshape2, win, name='delta')
    if use_blstm:
        f_lstm, b_lstm = create_blstm(l_delta, l_mask, lstm_size, cell_parameters, gate_parameters, 'blstm1',
                                      use_peepholes)
        l_sum1 = ElemwiseSumLayer([f_lstm, b_lstm], name='sum1')
        l_reshape3 = ReshapeLayer(l_sum1, (-1, lstm_size), name='reshape3')
    else:
        l_lstm = create_lstm(l_delta, l_mask, lstm_size, cell_parameters, gate_parameters, 'lstm', use_peepholes)
        l_reshape3 = ReshapeLayer(l_lstm, (-1, lstm_size), name='reshape3')
    l_softmax = DenseLayer(l_reshape3, num_units=output_classes, nonlinearity=softmax, name='softmax')
    l_out = ReshapeLayer(l_softmax, (-1, None, output_classes), name='output')
    return l_out


def extract_encoder_weights(network, names, saveas):
    """`modelzoo/deltanet_majority_vote.py:137-155`: {saveas[i][0]: W, saveas[i][1]: b} by layer name."""
    layers = get_all_layers(network)
    d = {}
    for i, name in enumerate(names):
        for l in layers:
            if l.name == name:
                d[saveas[i][0]] = l.W.get_value()
                d[saveas[i][1]] = l.b.get_value()
                break
    return d


def extract_lstm_weights(network, names, saveas):
    """`modelzoo/deltanet_majority_vote.py:158-196`: the 12 `{prefix}_w_*`/`{prefix}_b_*` keys per LSTM
    (peepholes and learned inits are dropped, as in the reference)."""
    layers = get_all_layers(network)
    d = {}
    for i, name in enumerate(names):
        for l in layers:
            if l.name == name:
                for g in ('cell', 'forgetgate', 'ingate', 'outgate'):
                    d['{}_w_hid_to_{}'.format(saveas[i], g)] = getattr(l, 'W_hid_to_' + g).get_value()
                    d['{}_w_in_to_{}'.format(saveas[i], g)] = getattr(l, 'W_in_to_' + g).get_value()
                    d['{}_b_{}'.format(saveas[i], g)] = getattr(l, 'b_' + g).get_value()
                break
    return d
