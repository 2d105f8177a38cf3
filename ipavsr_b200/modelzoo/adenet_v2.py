"""AdeNet v2: late fusion, frame-level head — mirrors `modelzoo/adenet_v2.py:12-94`.
Raw stream: Encoder -> Delta -> LSTM `lstm_bn`;  DCT stream: Delta -> LSTM `lstm_dct` (v2 *does* apply the
DeltaLayer to the DCT stream, :43);  fuse sum|adasum|concat;  aggregate BLSTM *without* peepholes (:77 calls
`create_blstm` without `use_peepholes`);  per-frame softmax."""
from .. import init
from ..layers import InputLayer, LSTMLayer, DenseLayer, ReshapeLayer, ElemwiseSumLayer, DeltaLayer
from ..nonlinearities import rectify, softmax
from ..custom.layers import create_blstm
from .pretrained_encoder import create_pretrained_encoder
from ._common import gates, fuse


def create_model(dbn, input_shape, input_var, mask_shape, mask_var, dct_shape, dct_var, lstm_size=250,
                 win=None, output_classes=26, fusiontype='sum', w_init_fn=init.GlorotUniform(),
                 use_peepholes=False, nonlinearities=rectify):
    weights, biases, shapes, nonlinearities = dbn      # the kwarg is dead in the reference too (:15,:17)
    names = ['fc1', 'fc2', 'fc3', 'bottleneck']
    gate_parameters, cell_parameters = gates(w_init_fn)
    l_in = InputLayer(input_shape, input_var, 'input')
    l_mask = InputLayer(mask_shape, mask_var, 'mask')
    l_dct = InputLayer(dct_shape, dct_var, 'dct')
    l_reshape1 = ReshapeLayer(l_in, (-1, input_shape[-1]), name='reshape1')
    l_encoder = create_pretrained_encoder(l_reshape1, weights, biases, shapes, nonlinearities, names)
    encoder_len = l_encoder.output_shape[-1]
    l_reshape2 = ReshapeLayer(l_encoder, (None, None, encoder_len), name='reshape2')
    l_delta = DeltaLayer(l_reshape2, win, name='delta')
    l_delta_dct = DeltaLayer(l_dct, win, name='delta_dct')

    def stream_lstm(incoming, name):
        return LSTMLayer(incoming, lstm_size, peepholes=use_peepholes, mask_input=l_mask,
                         ingate=gate_parameters, forgetgate=gate_parameters, cell=cell_parameters,
                         outgate=gate_parameters, learn_init=True, grad_clipping=5., name=name)

    l_lstm_bn = stream_lstm(l_delta, 'lstm_bn')
    l_lstm_dct = stream_lstm(l_delta_dct, 'lstm_dct')
    l_fuse = fuse(fusiontype, [l_lstm_bn, l_lstm_dct],
                  {'sum': 'sum1', 'adasum': 'adasum', 'concat': 'concat'}, strict=True)
    f_lstm_agg, b_lstm_agg = create_blstm(l_fuse, l_mask, lstm_size, cell_parameters, gate_parameters, 'lstm_agg')
    l_sum2 = ElemwiseSumLayer([f_lstm_agg, b_lstm_agg], name='sum2')
    l_reshape3 = ReshapeLayer(l_sum2, (-1, lstm_size), name='reshape3')
    l_softmax = DenseLayer(l_reshape3, num_units=output_classes, nonlinearity=softmax, name='softmax')
    l_out = ReshapeLayer(l_softmax, (-1, None, output_classes), name='output')
    return l_out, l_fuse
