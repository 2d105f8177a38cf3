"""Audio-visual net assembled from separately built sub-streams — mirrors `modelzoo/avnet.py:12-114`
(`cuave/audio_visual_runner.py:331-341`): `create_pretrained_substream` builds Input -> Encoder -> Delta -> LSTM for one
modality, `create_model` fuses a list of such sub-streams, BLSTM (no peepholes: `custom.layers.create_blstm` default),
per-frame softmax.

One deviation: the reference creates a new `InputLayer(mask_shape, mask_var, 'mask')` in every sub-stream and again in
`create_model` (:42, :87) — S+1 layer objects over the SAME Theano variable.  Here one mask InputLayer is kept per mask
variable (compiled functions bind one array per variable); layer names, parameter order and arithmetic are unaffected."""
from .. import init
from ..layers import InputLayer, LSTMLayer, DenseLayer, ReshapeLayer, ElemwiseSumLayer, DeltaLayer
from ..nonlinearities import rectify, linear, softmax
from ..custom.layers import create_blstm
from .pretrained_encoder import create_pretrained_encoder
from .adenet_v2_1 import extract_weights          # noqa: F401  (same body as `modelzoo/avnet.py:12-28`)
from ._common import gates, fuse


def _mask_layer(mask_shape, mask_var):
    l = getattr(mask_var, '_ipavsr_mask_layer', None)
    if l is None:
        l = InputLayer(mask_shape, mask_var, 'mask')
        mask_var._ipavsr_mask_layer = l
    return l


def create_pretrained_substream(weights, biases, input_shape, input_var, mask_shape, mask_var, name, lstm_size=250,
                                win=None, nonlinearity=rectify, w_init_fn=init.Orthogonal(), use_peepholes=True):
    gate_parameters, cell_parameters = gates(w_init_fn)
    l_input = InputLayer(input_shape, input_var, 'input_' + name)
    l_mask = _mask_layer(mask_shape, mask_var)
    l_reshape1_raw = ReshapeLayer(l_input, (-1, input_shape[-1]), name='reshape1_' + name)
    l_encoder_raw = create_pretrained_encoder(l_reshape1_raw, weights, biases, [2000, 1000, 500, 50],
                                              [nonlinearity, nonlinearity, nonlinearity, linear],
                                              ['fc1_' + name, 'fc2_' + name, 'fc3_' + name, 'bottleneck_' + name])
    l_reshape2 = ReshapeLayer(l_encoder_raw, (None, None, l_encoder_raw.output_shape[-1]), name='reshape2_' + name)
    l_delta = DeltaLayer(l_reshape2, win, name='delta_' + name)
    return LSTMLayer(l_delta, int(lstm_size), peepholes=use_peepholes, mask_input=l_mask, ingate=gate_parameters,
                     forgetgate=gate_parameters, cell=cell_parameters, outgate=gate_parameters, learn_init=True,
                     grad_clipping=5., name='lstm_' + name)


def create_model(substreams, mask_shape, mask_var, lstm_size=250, output_classes=26, fusiontype='concat',
                 w_init_fn=init.Orthogonal(), use_peepholes=True):
    gate_parameters, cell_parameters = gates(w_init_fn)
    l_mask = _mask_layer(mask_shape, mask_var)
    l_fuse = fuse(fusiontype, list(substreams), {'sum': 'sum1', 'adasum': 'adasum1', 'concat': 'concat'}, strict=False)
    f_lstm_agg, b_lstm_agg = create_blstm(l_fuse, l_mask, lstm_size, cell_parameters, gate_parameters, 'lstm_agg')
    l_sum2 = ElemwiseSumLayer([f_lstm_agg, b_lstm_agg], name='sum2')
    l_reshape3 = ReshapeLayer(l_sum2, (-1, lstm_size), name='reshape3')
    l_softmax = DenseLayer(l_reshape3, num_units=output_classes, nonlinearity=softmax, name='softmax')
    l_out = ReshapeLayer(l_softmax, (-1, None, output_classes), name='output')
    return l_out, l_fuse
