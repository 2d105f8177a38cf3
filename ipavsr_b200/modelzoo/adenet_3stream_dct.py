"""3-stream AdeNet whose third stream has no encoder (DCT features -> Delta -> LSTM) — mirrors
`modelzoo/adenet_3stream_dct.py:12-119`."""
from .. import init
from . import _nstream


def create_model(s1_ae, s2_ae, s1_shape, s1_var, s2_shape, s2_var, s3_shape, s3_var, mask_shape, mask_var,
                 lstm_size=250, win=None, output_classes=26, fusiontype='concat', w_init_fn=init.Orthogonal(),
                 use_peepholes=True):
    return _nstream.build([s1_ae, s2_ae, None], [s1_shape, s2_shape, s3_shape], [s1_var, s2_var, s3_var], mask_shape,
                          mask_var, lstm_size, win, output_classes, fusiontype, w_init_fn, use_peepholes)
