"""Delta + (B)LSTM on precomputed features, frame-level head — mirrors `modelzoo/deltanet_v1.py:8-42`.
Note `window` comes *before* `lstm_size` in this builder."""
from .. import init
from ..layers import InputLayer, DenseLayer, ReshapeLayer, ElemwiseSumLayer, DeltaLayer
from ..nonlinearities import softmax
from ..custom.layers import create_blstm, create_lstm
from ._common import gates


def create_model(input_shape, input_var, mask_shape, mask_var, window, lstm_size=250, output_classes=26,
                 w_init=init.GlorotUniform(), use_peepholes=False, use_blstm=True):
    gate_parameters, cell_parameters = gates(w_init)
    l_in = InputLayer(input_shape, input_var, 'input')
    l_mask = InputLayer(mask_shape, mask_var, name='mask')
    l_delta = DeltaLayer(l_in, window, name='delta')
    if use_blstm:
        f_lstm, b_lstm = create_blstm(l_delta, l_mask, lstm_size, cell_parameters, gate_parameters, 'lstm',
                                      use_peepholes)
        l_sum = ElemwiseSumLayer([f_lstm, b_lstm], name='sum')
        l_reshape = ReshapeLayer(l_sum, (-1, lstm_size), name='reshape')
    else:
        l_lstm = create_lstm(l_delta, l_mask, lstm_size, cell_parameters, gate_parameters, 'lstm', use_peepholes)
        l_reshape = ReshapeLayer(l_lstm, (-1, lstm_size), name='reshape')
    l_softmax = DenseLayer(l_reshape, num_units=output_classes, nonlinearity=softmax, name='softmax')
    l_out = ReshapeLayer(l_softmax, (-1, None, output_classes), name='output')
    return l_out
