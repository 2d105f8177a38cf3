"""AdeNet v3 (README-era trimodal: raw + DCT + diff) — mirrors `modelzoo/adenet_v3.py:12-188`.
Sigmoid encoders (names `*_raw`, `*_diff`), dropout p=.5 (.2 on DCT), stream LSTMs `int(lstm_size/(1-0.5))`
wide with peepholes (Lasagne default), no DeltaLayer on DCT, fuse, dropout, BLSTM(2*lstm_size),
slice T-1, softmax."""
from .. import init
from ..layers import (InputLayer, LSTMLayer, DenseLayer, SliceLayer, ReshapeLayer, ElemwiseSumLayer,
                      DropoutLayer, DeltaLayer)
from ..nonlinearities import sigmoid, linear, softmax
from .lstm_classifier_baseline import create_blstm
from .pretrained_encoder import extract_dbn_weights
from ._common import gates, fuse


def create_pretrained_encoder(weights, biases, names, incoming):
    l_1 = DenseLayer(incoming, 2000, W=weights[0], b=biases[0], nonlinearity=sigmoid, name=names[0])
    l_2 = DenseLayer(l_1, 1000, W=weights[1], b=biases[1], nonlinearity=sigmoid, name=names[1])
    l_3 = DenseLayer(l_2, 500, W=weights[2], b=biases[2], nonlinearity=sigmoid, name=names[2])
    l_4 = DenseLayer(l_3, 50, W=weights[3], b=biases[3], nonlinearity=linear, name=names[3])
    return l_4


extract_weights = extract_dbn_weights


def create_model(ae, diff_ae, input_shape, input_var, mask_shape, mask_var, dct_shape, dct_var,
                 diff_shape, diff_var, lstm_size=250, win=None, output_classes=26, fusiontype='sum'):
    return _build(ae, diff_ae, input_shape, input_var, mask_shape, mask_var, dct_shape, dct_var, diff_shape, diff_var,
                  lstm_size, win, output_classes, fusiontype)


def _build(ae, diff_ae, input_shape, input_var, mask_shape, mask_var, dct_shape, dct_var, diff_shape, diff_var,
           lstm_size, win, output_classes, fusiontype):
    """dct_shape=None drops the DCT stream (`modelzoo/adenet_v6.py:64-175`)."""
    bn_weights, bn_biases = extract_weights(ae)
    diff_weights, diff_biases = extract_weights(diff_ae)
    gate_parameters, cell_parameters = gates(init.Orthogonal())
    l_raw = InputLayer(input_shape, input_var, 'raw_im')
    l_mask = InputLayer(mask_shape, mask_var, 'mask')
    l_dct = InputLayer(dct_shape, dct_var, 'dct') if dct_shape is not None else None
    l_diff = InputLayer(diff_shape, diff_var, 'diff_im')

    l_reshape1_raw = ReshapeLayer(l_raw, (-1, input_shape[-1]), name='reshape1_raw')
    l_encoder_raw = create_pretrained_encoder(bn_weights, bn_biases,
                                              ['fc1_raw', 'fc2_raw', 'fc3_raw', 'bottleneck_raw'], l_reshape1_raw)
    raw_len = l_encoder_raw.output_shape[-1]
    l_reshape2_raw = ReshapeLayer(l_encoder_raw, (None, None, raw_len), name='reshape2_raw')
    l_delta_raw = DeltaLayer(l_reshape2_raw, win, name='delta_raw')

    l_reshape1_diff = ReshapeLayer(l_diff, (-1, diff_shape[-1]), name='reshape1_diff')
    l_encoder_diff = create_pretrained_encoder(diff_weights, diff_biases,
                                               ['fc1_diff', 'fc2_diff', 'fc3_diff', 'bottleneck_diff'],
                                               l_reshape1_diff)
    diff_len = l_encoder_diff.output_shape[-1]
    l_reshape2_diff = ReshapeLayer(l_encoder_diff, (None, None, diff_len), name='reshape2_diff')
    l_delta_diff = DeltaLayer(l_reshape2_diff, win, name='delta_diff')

    def stream_lstm(incoming, name):
        return LSTMLayer(incoming, int(lstm_size / (1 - 0.5)), mask_input=l_mask, ingate=gate_parameters,
                         forgetgate=gate_parameters, cell=cell_parameters, outgate=gate_parameters,
                         learn_init=True, grad_clipping=5., name=name)

    l_delta_raw_drop = DropoutLayer(l_delta_raw, name='dropout_raw')
    l_lstm_raw = stream_lstm(l_delta_raw_drop, 'lstm_raw')
    streams = [l_lstm_raw]
    if l_dct is not None:
        l_dct_drop = DropoutLayer(l_dct, p=0.2, name='dropout_dct')
        streams.append(stream_lstm(l_dct_drop, 'lstm_dct'))
    l_delta_diff_drop = DropoutLayer(l_delta_diff, name='dropout_diff')
    streams.append(stream_lstm(l_delta_diff_drop, 'lstm_diff'))

    l_fuse = fuse(fusiontype, streams,
                  {'sum': 'sum1', 'adasum': 'adasum1', 'concat': 'concat'}, strict=False)
    l_drop_agg = DropoutLayer(l_fuse, name='dropout_agg')
    f_lstm_agg, b_lstm_agg = create_blstm(l_drop_agg, l_mask, lstm_size * 2, cell_parameters, gate_parameters,
                                          'lstm_agg')
    l_sum2 = ElemwiseSumLayer([f_lstm_agg, b_lstm_agg], name='sum2')
    l_forward_slice1 = SliceLayer(l_sum2, -1, 1, name='slice1')
    l_out = DenseLayer(l_forward_slice1, num_units=output_classes, nonlinearity=softmax, name='output')
    return l_out, l_fuse
