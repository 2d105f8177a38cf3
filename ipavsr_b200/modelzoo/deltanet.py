"""DeltaNet, sequence-level head — mirrors `modelzoo/deltanet.py:12-76`."""
from .. import init
from ..layers import InputLayer, DenseLayer, SliceLayer, ReshapeLayer, ElemwiseSumLayer, DeltaLayer
from ..nonlinearities import linear, rectify, softmax
from ..custom.layers import create_blstm
from .pretrained_encoder import create_pretrained_encoder, extract_dbn_weights
from ._common import gates


def create_model_using_pretrained_encoder(weights, biases, input_shape, input_var, mask_shape, mask_var,
                                          lstm_size=250, win=None, output_classes=26,
                                          w_init_fn=init.Orthogonal(), use_peepholes=False,
                                          nonlinearities=rectify):
    gate_parameters, cell_parameters = gates(w_init_fn)
    l_in = InputLayer(input_shape, input_var, 'input')
    l_mask = InputLayer(mask_shape, mask_var, 'mask')
    l_reshape1 = ReshapeLayer(l_in, (-1, input_shape[-1]), name='reshape1')
    l_encoder = create_pretrained_encoder(l_reshape1, weights, biases, [2000, 1000, 500, 50],
                                          [nonlinearities, nonlinearities, nonlinearities, linear],
                                          ['fc1', 'fc2', 'fc3', 'bottleneck'])
    encoder_len = l_encoder.output_shape[-1]
    l_reshape2 = ReshapeLayer(l_encoder, (None, None, encoder_len), name='reshape2')
    l_delta = DeltaLayer(l_reshape2, win, name='delta')
    l_lstm, l_lstm_back = create_blstm(l_delta, l_mask, lstm_size, cell_parameters, gate_parameters, 'bstm1',
                                       use_peepholes)
    l_sum1 = ElemwiseSumLayer([l_lstm, l_lstm_back], name='sum1')
    l_forward_slice1 = SliceLayer(l_sum1, -1, 1, name='slice1')
    l_out = DenseLayer(l_forward_slice1, num_units=output_classes, nonlinearity=softmax, name='output')
    return l_out


def create_model(dbn, input_shape, input_var, mask_shape, mask_var, lstm_size=250, win=None,
                 output_classes=26):
    weights, biases = extract_dbn_weights(dbn)
    return create_model_using_pretrained_encoder(weights, biases, input_shape, input_var, mask_shape, mask_var,
                                                 lstm_size, win, output_classes)
