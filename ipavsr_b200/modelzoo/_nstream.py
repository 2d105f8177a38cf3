"""Shared wiring of the N-stream late-fusion builders (`adenet_3stream.py:145-262`, `adenet_4stream.py:12-159`):
every stream is Encoder -> Delta -> LSTM(p=use_peepholes); the aggregate BLSTM never gets peepholes."""
from ..layers import InputLayer, LSTMLayer, DenseLayer, ReshapeLayer, ElemwiseSumLayer, DeltaLayer
from ..nonlinearities import softmax
from ..custom.layers import create_blstm
from .pretrained_encoder import create_pretrained_encoder
from ._common import gates, fuse


def build(aes, shapes, variables, mask_shape, mask_var, lstm_size, win, output_classes, fusiontype, w_init_fn,
          use_peepholes):
    gate_parameters, cell_parameters = gates(w_init_fn)
    S = len(aes)
    # InputLayer creation order of the reference: s1, mask, s2, s3[, s4]
    l_ins = [None] * S
    l_ins[0] = InputLayer(shapes[0], variables[0], 's1_im')
    l_mask = InputLayer(mask_shape, mask_var, 'mask')
    for k in range(1, S):
        l_ins[k] = InputLayer(shapes[k], variables[k], 's%d_im' % (k + 1))
    deltas = []
    for k in range(S):
        s = 's%d' % (k + 1)
        weights, biases, enc_shapes, nonlins = aes[k]
        l_r1 = ReshapeLayer(l_ins[k], (-1, shapes[k][-1]), name='reshape1_' + s)
        l_enc = create_pretrained_encoder(l_r1, weights, biases, enc_shapes, nonlins,
                                          ['fc1_' + s, 'fc2_' + s, 'fc3_' + s, 'bottleneck_' + s])
        l_r2 = ReshapeLayer(l_enc, (None, None, l_enc.output_shape[-1]), name='reshape2_' + s)
        deltas.append(DeltaLayer(l_r2, win, name='delta_' + s))
    lstms = []
    for k in range(S):
        lstms.append(LSTMLayer(deltas[k], int(lstm_size), peepholes=use_peepholes, mask_input=l_mask,
                               ingate=gate_parameters, forgetgate=gate_parameters, cell=cell_parameters,
                               outgate=gate_parameters, learn_init=True, grad_clipping=5.,
                               name='lstm_s%d' % (k + 1)))
    l_fuse = fuse(fusiontype, lstms, {'sum': 'sum1', 'adasum': 'adasum1', 'concat': 'concat'}, strict=False)
    f_lstm_agg, b_lstm_agg = create_blstm(l_fuse, l_mask, lstm_size, cell_parameters, gate_parameters, 'lstm_agg')
    l_sum2 = ElemwiseSumLayer([f_lstm_agg, b_lstm_agg], name='sum2')
    l_reshape3 = ReshapeLayer(l_sum2, (-1, lstm_size), name='reshape3')
    l_softmax = DenseLayer(l_reshape3, num_units=output_classes, nonlinearity=softmax, name='softmax')
    l_out = ReshapeLayer(l_softmax, (-1, None, output_classes), name='output')
    return l_out, l_fuse
