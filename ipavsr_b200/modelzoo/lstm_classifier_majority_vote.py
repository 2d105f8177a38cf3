"""LSTM / BLSTM frame-level classifier on one feature stream (the majority-vote baseline) — mirrors
`modelzoo/lstm_classifier_majority_vote.py:10-43`."""
from .. import init
from ..layers import InputLayer, DenseLayer, ReshapeLayer, ElemwiseSumLayer
from ..nonlinearities import softmax
from ..custom.layers import create_blstm, create_lstm
from ._common import gates


def create_model(input_shape, input_var, mask_shape, mask_var, lstm_size=250, output_classes=26,
                 w_init=init.GlorotUniform(), use_peepholes=False, use_blstm=True):
    gate_parameters, cell_parameters = gates(w_init)
    l_in = InputLayer(input_shape, input_var, 'input')
    l_mask = InputLayer(mask_shape, mask_var, 'mask')
    if use_blstm:
        f_lstm, b_lstm = create_blstm(l_in, l_mask, lstm_size, cell_parameters, gate_parameters, 'lstm', use_peepholes)
        l_sum = ElemwiseSumLayer([f_lstm, b_lstm], name='sum')
        l_reshape = ReshapeLayer(l_sum, (-1, lstm_size), name='reshape')
    else:
        l_lstm = create_lstm(l_in, l_mask, lstm_size, cell_parameters, gate_parameters, 'lstm', use_peepholes)
        l_reshape = ReshapeLayer(l_lstm, (-1, lstm_size), name='reshape')
    l_softmax = DenseLayer(l_reshape, num_units=output_classes, nonlinearity=softmax, name='softmax')
    l_out = ReshapeLayer(l_softmax, (-1, None, output_classes), name='output')
    return l_out
