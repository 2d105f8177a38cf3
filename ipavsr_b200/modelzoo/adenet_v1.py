"""AdeNet v1: early (feature) fusion — mirrors `modelzoo/adenet_v1.py:11-109`.
Encoder(sigmoid) -> BatchNorm(bottleneck) -> Delta ++ DCT -> BLSTM(H) -> BLSTM(2H) -> slice T-1 -> softmax.
Peepholes are ON (kwarg omitted in the reference => Lasagne default)."""
from .. import init
from ..layers import (InputLayer, DenseLayer, ConcatLayer, SliceLayer, ReshapeLayer, ElemwiseSumLayer,
                      BatchNormLayer, DeltaLayer)
from ..nonlinearities import sigmoid, linear, softmax
from .lstm_classifier_baseline import create_blstm
from .pretrained_encoder import extract_dbn_weights
from ._common import gates


def create_pretrained_encoder(weights, biases, incoming):
    l_1 = DenseLayer(incoming, 2000, W=weights[0], b=biases[0], nonlinearity=sigmoid, name='fc1')
    l_2 = DenseLayer(l_1, 1000, W=weights[1], b=biases[1], nonlinearity=sigmoid, name='fc2')
    l_3 = DenseLayer(l_2, 500, W=weights[2], b=biases[2], nonlinearity=sigmoid, name='fc3')
    l_4 = DenseLayer(l_3, 50, W=weights[3], b=biases[3], nonlinearity=linear, name='bottleneck')
    return l_4


def create_model(dbn, input_shape, input_var, mask_shape, mask_var, dct_shape, dct_var, lstm_size=250,
                 win=None, output_classes=26):
    return _build(dbn, input_shape, input_var, mask_shape, mask_var, dct_shape, dct_var, lstm_size, win, output_classes,
                  False)


def _build(dbn, input_shape, input_var, mask_shape, mask_var, dct_shape, dct_var, lstm_size, win, output_classes, dropout):
    """dropout=True is `modelzoo/adenet_v1_1.py:47-104`: dropout before each BLSTM, the first BLSTM 2*lstm_size wide."""
    from ..layers import DropoutLayer
    weights, biases = extract_dbn_weights(dbn)
    gate_parameters, cell_parameters = gates(init.Orthogonal())
    l_in = InputLayer(input_shape, input_var, 'input')
    l_mask = InputLayer(mask_shape, mask_var, 'mask')
    l_dct = InputLayer(dct_shape, dct_var, 'dct')
    l_reshape1 = ReshapeLayer(l_in, (-1, input_shape[-1]), name='reshape1')
    l_encoder = create_pretrained_encoder(weights, biases, l_reshape1)
    l_encoder_bn = BatchNormLayer(l_encoder, name='batchnorm1')
    encoder_len = l_encoder.output_shape[-1]
    l_reshape2 = ReshapeLayer(l_encoder_bn, (None, None, encoder_len), name='reshape2')
    l_delta = DeltaLayer(l_reshape2, win, name='delta')
    l_concat = ConcatLayer([l_delta, l_dct], axis=2, name='concat')
    l_top = DropoutLayer(l_concat, name='dropout1') if dropout else l_concat
    l_lstm, l_lstm_back = create_blstm(l_top, l_mask, lstm_size * 2 if dropout else lstm_size, cell_parameters,
                                       gate_parameters, 'lstm1')
    l_sum1 = ElemwiseSumLayer([l_lstm, l_lstm_back], name='sum1')
    l_top = DropoutLayer(l_sum1, name='dropout2') if dropout else l_sum1
    l_lstm2, l_lstm2_back = create_blstm(l_top, l_mask, lstm_size * 2, cell_parameters, gate_parameters,
                                         'lstm2')
    l_sum2 = ElemwiseSumLayer([l_lstm2, l_lstm2_back])
    l_forward_slice1 = SliceLayer(l_sum2, -1, 1, name='slice1')
    l_out = DenseLayer(l_forward_slice1, num_units=output_classes, nonlinearity=softmax, name='output')
    return l_out, l_concat
