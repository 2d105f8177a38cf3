"""AdeNet v4: raw (sigmoid encoder -> Delta -> dropout) + DCT (dropout .2) LSTMs 2*lstm_size wide with peepholes
(Lasagne default), summed, dropout, ONE forward LSTM 2*lstm_size, slice T-1, softmax — mirrors `modelzoo/adenet_v4.py:48-145`;
returns (l_out, l_sum1)."""
from .. import init
from ..layers import (InputLayer, LSTMLayer, DenseLayer, SliceLayer, ReshapeLayer, ElemwiseSumLayer, DropoutLayer,
                      DeltaLayer)
from ..nonlinearities import softmax
from .adenet_v1 import create_pretrained_encoder
from .pretrained_encoder import extract_dbn_weights
from ._common import gates


def create_model(dbn, input_shape, input_var, mask_shape, mask_var, dct_shape, dct_var, lstm_size=250, win=None,
                 output_classes=26):
    weights, biases = extract_dbn_weights(dbn)
    gate_parameters, cell_parameters = gates(init.Orthogonal())
    l_in = InputLayer(input_shape, input_var, 'input')
    l_mask = InputLayer(mask_shape, mask_var, 'mask')
    l_dct = InputLayer(dct_shape, dct_var, 'dct')
    l_reshape1 = ReshapeLayer(l_in, (-1, input_shape[-1]), name='reshape1')
    l_encoder = create_pretrained_encoder(weights, biases, l_reshape1)
    l_reshape2 = ReshapeLayer(l_encoder, (None, None, l_encoder.output_shape[-1]), name='reshape2')
    l_delta = DeltaLayer(l_reshape2, win, name='delta')
    l_delta_drop = DropoutLayer(l_delta, name='dropout_delta')
    l_dct_drop = DropoutLayer(l_dct, p=0.2, name='dropout_dct')

    def lstm(incoming, name):
        return LSTMLayer(incoming, lstm_size * 2, mask_input=l_mask, ingate=gate_parameters, forgetgate=gate_parameters,
                         cell=cell_parameters, outgate=gate_parameters, learn_init=True, grad_clipping=5., name=name)

    l_lstm_bn = lstm(l_delta_drop, 'lstm_bn')
    l_lstm_dct = lstm(l_dct_drop, 'lstm_dct')
    l_sum1 = ElemwiseSumLayer([l_lstm_bn, l_lstm_dct], name='sum1')
    l_sum1_drop = DropoutLayer(l_sum1, name='dropout_agg')
    l_lstm_agg = lstm(l_sum1_drop, 'lstm_agg')
    l_forward_slice1 = SliceLayer(l_lstm_agg, -1, 1, name='slice1')
    l_out = DenseLayer(l_forward_slice1, num_units=output_classes, nonlinearity=softmax, name='output')
    return l_out, l_sum1
