"""AdeNet v2.2: two encoder streams (4-tuples), late fusion, frame-level head — mirrors `modelzoo/adenet_v2_2.py:41-130`.
The file-local create_blstm defaults to use_peepholes=True (:13), so the aggregate BLSTM has peepholes."""
from .. import init
from . import _nstream


def create_model(ae, s2_ae, input_shape, input_var, mask_shape, mask_var, s2_shape, s2_var, lstm_size=250, win=None,
                 output_classes=26, fusiontype='concat', w_init_fn=init.Orthogonal(), use_peepholes=True):
    return _nstream.build([ae, s2_ae], [input_shape, s2_shape], [input_var, s2_var], mask_shape, mask_var, lstm_size,
                          win, output_classes, fusiontype, w_init_fn, use_peepholes, agg_peepholes=True)
