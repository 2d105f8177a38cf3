"""AdeNet v2 without DeltaLayers (encoder bottlenecks feed the LSTMs directly; no `win` argument) — mirrors
`modelzoo/adenet_v2_nodelta.py:40-128`; aggregate BLSTM with peepholes (file-local create_blstm, :12)."""
from .. import init
from . import _nstream


def create_model(ae, s2_ae, input_shape, input_var, mask_shape, mask_var, s2_shape, s2_var, lstm_size=250,
                 output_classes=26, fusiontype='concat', w_init_fn=init.Orthogonal(), use_peepholes=True):
    return _nstream.build([ae, s2_ae], [input_shape, s2_shape], [input_var, s2_var], mask_shape, mask_var, lstm_size,
                          None, output_classes, fusiontype, w_init_fn, use_peepholes, agg_peepholes=True, delta=False)
