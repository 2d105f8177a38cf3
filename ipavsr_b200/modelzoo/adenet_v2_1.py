"""AdeNet v2.1: raw + diff-image encoders (legacy auto-encoder objects, rectify), Delta, LSTMs, fusion, BLSTM with
peepholes (file-local create_blstm default, :14), sequence-level head (slice T-1) — mirrors `modelzoo/adenet_v2_1.py:63-171`."""
from .. import init
from ..layers import InputLayer, LSTMLayer, DenseLayer, SliceLayer, ReshapeLayer, ElemwiseSumLayer, DeltaLayer
from ..nonlinearities import rectify, linear, softmax
from ..custom.layers import create_blstm, create_lstm
from .pretrained_encoder import create_pretrained_encoder, extract_dbn_weights
from ._common import gates, fuse


def extract_weights(ae):
    """`modelzoo/adenet_v2_1.py:44-60`."""
    weights, biases = extract_dbn_weights(ae)
    return weights, biases, [2000, 1000, 500, 50], [rectify, rectify, rectify, linear]


def build_raw_diff(ae4, diff4, input_shape, input_var, mask_shape, mask_var, diff_shape, diff_var, lstm_size, win,
                   output_classes, fusiontype, w_init_fn, use_peepholes, head):
    """Shared by v2_1 (head='sequence': BLSTM, slice, softmax 'output') and v2_4 (head='frame_lstm': one forward LSTM
    `f_lstm_agg`, unnamed reshape, per-frame softmax)."""
    bn_weights, bn_biases, bn_shapes, bn_nonlinearities = ae4
    diff_weights, diff_biases, diff_shapes, diff_nonlinearities = diff4
    gate_parameters, cell_parameters = gates(w_init_fn)
    l_raw = InputLayer(input_shape, input_var, 'raw_im')
    l_mask = InputLayer(mask_shape, mask_var, 'mask')
    l_diff = InputLayer(diff_shape, diff_var, 'diff_im')

    def stream(l_in, shape, weights, biases, shapes, nonlins, s, units):
        l_r1 = ReshapeLayer(l_in, (-1, shape[-1]), name='reshape1_' + s)
        l_enc = create_pretrained_encoder(l_r1, weights, biases, shapes, nonlins,
                                          ['fc1_' + s, 'fc2_' + s, 'fc3_' + s, 'bottleneck_' + s])
        l_r2 = ReshapeLayer(l_enc, (None, None, l_enc.output_shape[-1]), name='reshape2_' + s)
        l_delta = DeltaLayer(l_r2, win, name='delta_' + s)
        return LSTMLayer(l_delta, units, peepholes=use_peepholes, mask_input=l_mask, ingate=gate_parameters,
                         forgetgate=gate_parameters, cell=cell_parameters, outgate=gate_parameters, learn_init=True,
                         grad_clipping=5., name='lstm_' + s)

    l_lstm_raw = stream(l_raw, input_shape, bn_weights, bn_biases, bn_shapes, bn_nonlinearities, 'raw', int(lstm_size))
    l_lstm_diff = stream(l_diff, diff_shape, diff_weights, diff_biases, diff_shapes, diff_nonlinearities, 'diff',
                         lstm_size)
    l_fuse = fuse(fusiontype, [l_lstm_raw, l_lstm_diff], {'sum': 'sum1', 'adasum': 'adasum1', 'concat': 'concat'},
                  strict=False)
    if head == 'sequence':
        f_lstm_agg, b_lstm_agg = create_blstm(l_fuse, l_mask, lstm_size, cell_parameters, gate_parameters, 'lstm_agg',
                                              True)
        l_sum2 = ElemwiseSumLayer([f_lstm_agg, b_lstm_agg], name='sum2')
        l_forward_slice1 = SliceLayer(l_sum2, -1, 1, name='slice1')
        l_out = DenseLayer(l_forward_slice1, num_units=output_classes, nonlinearity=softmax, name='output')
        return l_out, l_fuse
    f_lstm_agg = create_lstm(l_fuse, l_mask, lstm_size, cell_parameters, gate_parameters, 'f_lstm_agg', True)
    l_reshape3 = ReshapeLayer(f_lstm_agg, (-1, lstm_size))
    l_softmax = DenseLayer(l_reshape3, num_units=output_classes, nonlinearity=softmax, name='softmax')
    l_out = ReshapeLayer(l_softmax, (-1, None, output_classes), name='output')
    return l_out, l_fuse


def create_model(ae, diff_ae, input_shape, input_var, mask_shape, mask_var, diff_shape, diff_var, lstm_size=250,
                 win=None, output_classes=26, fusiontype='concat', w_init_fn=init.Orthogonal(), use_peepholes=True):
    return build_raw_diff(extract_weights(ae), extract_weights(diff_ae), input_shape, input_var, mask_shape, mask_var,
                          diff_shape, diff_var, lstm_size, win, output_classes, fusiontype, w_init_fn, use_peepholes,
                          'sequence')
