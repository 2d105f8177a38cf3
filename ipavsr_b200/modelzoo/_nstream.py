"""Shared wiring of the N-stream late-fusion builders (`adenet_3stream.py:145-262`, `adenet_4stream.py:12-159`):
every stream is Encoder -> Delta -> LSTM(p=use_peepholes); the aggregate BLSTM never gets peepholes."""
from ..layers import InputLayer, LSTMLayer, DenseLayer, ReshapeLayer, ElemwiseSumLayer, DeltaLayer
from ..nonlinearities import softmax
from ..custom.layers import create_blstm
from .pretrained_encoder import create_pretrained_encoder
from ._common import gates, fuse


def build(aes, shapes, variables, mask_shape, mask_var, lstm_size, win, output_classes, fusiontype, w_init_fn,
          use_peepholes, agg_peepholes=False, delta=True, stream_units=None, dropout=False, lstm_weights=None,
          use_blstm_substream=False):
    """Options cover the variants of the same wiring: `agg_peepholes` (files with a local create_blstm whose default is
    use_peepholes=True: `adenet_v2_2.py:13`, `adenet_v2_nodelta.py:12`), `delta=False` (`adenet_v2_nodelta.py:76-87`),
    `aes[k] is None` = a stream without an encoder (the DCT stream of `adenet_3stream_dct.py:82`), `dropout` + `stream_units`
    (`adenet_3stream_dropout.py:68-123`: dropout after every delta and after the fusion, LSTMs 2*lstm_size wide),
    `lstm_weights` (per-stream LSTM .mat dicts: `create_pretrained_model`, `adenet_2stream.py:12-114`,
    `adenet_3stream.py:12-142`, optionally forward+backward sub-stream LSTMs summed)."""
    from ..layers import DropoutLayer
    from ..custom.layers import create_pretrained_lstm
    gate_parameters, cell_parameters = gates(w_init_fn)
    S = len(aes)
    units = int(stream_units if stream_units is not None else lstm_size)
    # InputLayer creation order of the reference: s1, mask, s2, s3[, s4]
    l_ins = [None] * S
    l_ins[0] = InputLayer(shapes[0], variables[0], 's1_im')
    l_mask = InputLayer(mask_shape, mask_var, 'mask')
    for k in range(1, S):
        l_ins[k] = InputLayer(shapes[k], variables[k], 's%d_im' % (k + 1))
    deltas = []
    for k in range(S):
        s = 's%d' % (k + 1)
        if aes[k] is None:
            top = l_ins[k]
        else:
            weights, biases, enc_shapes, nonlins = aes[k]
            l_r1 = ReshapeLayer(l_ins[k], (-1, shapes[k][-1]), name='reshape1_' + s)
            l_enc = create_pretrained_encoder(l_r1, weights, biases, enc_shapes, nonlins,
                                              ['fc1_' + s, 'fc2_' + s, 'fc3_' + s, 'bottleneck_' + s])
            top = ReshapeLayer(l_enc, (None, None, l_enc.output_shape[-1]), name='reshape2_' + s)
        if delta:
            top = DeltaLayer(top, win, name='delta_' + s)
        if dropout:
            top = DropoutLayer(top, name='dropout_' + s)
        deltas.append(top)
    lstms = []
    for k in range(S):
        s = 's%d' % (k + 1)
        if lstm_weights is not None:
            f = create_pretrained_lstm(lstm_weights[k], 'f_lstm', deltas[k], l_mask, lstm_size, cell_parameters,
                                       gate_parameters, 'f_lstm_' + s, use_peepholes)
            if use_blstm_substream:
                b = create_pretrained_lstm(lstm_weights[k], 'b_lstm', deltas[k], l_mask, lstm_size, cell_parameters,
                                           gate_parameters, 'b_lstm_' + s, use_peepholes, backwards=True)
                f = ElemwiseSumLayer([f, b], name='sum_b_lstm_' + s)
            lstms.append(f)
            continue
        lstms.append(LSTMLayer(deltas[k], units, peepholes=use_peepholes, mask_input=l_mask,
                               ingate=gate_parameters, forgetgate=gate_parameters, cell=cell_parameters,
                               outgate=gate_parameters, learn_init=True, grad_clipping=5.,
                               name='lstm_' + s))
    l_fuse = fuse(fusiontype, lstms, {'sum': 'sum1', 'adasum': 'adasum1', 'concat': 'concat'}, strict=False)
    agg_in = DropoutLayer(l_fuse, name='concat_dropout') if dropout else l_fuse
    lstm_size = units if dropout else lstm_size              # adenet_3stream_dropout: the aggregate is lstm_size*2 wide too
    f_lstm_agg, b_lstm_agg = create_blstm(agg_in, l_mask, lstm_size, cell_parameters, gate_parameters, 'lstm_agg',
                                          agg_peepholes)
    l_sum2 = ElemwiseSumLayer([f_lstm_agg, b_lstm_agg], name='sum2')
    l_reshape3 = ReshapeLayer(l_sum2, (-1, lstm_size), name='reshape3')
    l_softmax = DenseLayer(l_reshape3, num_units=output_classes, nonlinearity=softmax, name='softmax')
    l_out = ReshapeLayer(l_softmax, (-1, None, output_classes), name='output')
    return l_out, l_fuse
