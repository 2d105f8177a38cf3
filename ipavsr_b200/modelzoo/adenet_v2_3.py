"""AdeNet v2.3: raw (legacy DBN object, rectify encoder -> Delta -> LSTM `lstm_bn`) + DCT (no DeltaLayer -> LSTM
`lstm_dct`), fusion, ONE forward aggregate LSTM `f_lstm_agg` with peepholes, frame-level head — mirrors
`modelzoo/adenet_v2_3.py:61-147`.  The gates' W_in initialiser is Orthogonal whatever `w_init_fn` is (:83)."""
from .. import init
from ..layers import InputLayer, LSTMLayer, DenseLayer, ReshapeLayer, DeltaLayer, Gate
from ..nonlinearities import rectify, linear, softmax, tanh
from ..custom.layers import create_lstm
from .pretrained_encoder import create_pretrained_encoder, extract_dbn_weights
from ._common import fuse


def create_model(dbn, input_shape, input_var, mask_shape, mask_var, dct_shape, dct_var, lstm_size=250, win=None,
                 output_classes=26, fusiontype='sum', w_init_fn=init.Orthogonal(), use_peepholes=True):
    weights, biases = extract_dbn_weights(dbn)
    shapes = [2000, 1000, 500, 50]
    nonlinearities = [rectify, rectify, rectify, linear]
    gate_parameters = Gate(W_in=init.Orthogonal(), W_hid=w_init_fn, b=init.Constant(0.))
    cell_parameters = Gate(W_in=w_init_fn, W_hid=w_init_fn, W_cell=None, b=init.Constant(0.), nonlinearity=tanh)
    l_in = InputLayer(input_shape, input_var, 'input')
    l_mask = InputLayer(mask_shape, mask_var, 'mask')
    l_dct = InputLayer(dct_shape, dct_var, 'dct')
    l_reshape1 = ReshapeLayer(l_in, (-1, input_shape[-1]), name='reshape1')
    l_encoder = create_pretrained_encoder(l_reshape1, weights, biases, shapes, nonlinearities,
                                          ['fc1', 'fc2', 'fc3', 'bottleneck'])
    l_reshape2 = ReshapeLayer(l_encoder, (None, None, l_encoder.output_shape[-1]), name='reshape2')
    l_delta = DeltaLayer(l_reshape2, win, name='delta')

    def lstm(incoming, name):
        return LSTMLayer(incoming, lstm_size, peepholes=use_peepholes, mask_input=l_mask, ingate=gate_parameters,
                         forgetgate=gate_parameters, cell=cell_parameters, outgate=gate_parameters, learn_init=True,
                         grad_clipping=5., name=name)

    l_lstm_bn = lstm(l_delta, 'lstm_bn')
    l_lstm_dct = lstm(l_dct, 'lstm_dct')
    l_fuse = fuse(fusiontype, [l_lstm_bn, l_lstm_dct], {'sum': 'sum1', 'adasum': 'adasum', 'concat': 'concat'},
                  strict=False)
    f_lstm_agg = create_lstm(l_fuse, l_mask, lstm_size, cell_parameters, gate_parameters, 'f_lstm_agg', True)
    l_reshape3 = ReshapeLayer(f_lstm_agg, (-1, lstm_size))
    l_softmax = DenseLayer(l_reshape3, num_units=output_classes, nonlinearity=softmax, name='softmax')
    l_out = ReshapeLayer(l_softmax, (-1, None, output_classes), name='output')
    return l_out, l_fuse
