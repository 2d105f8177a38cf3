"""AdeNet v2.4: raw + diff-image encoders (4-tuples), late fusion, ONE forward aggregate LSTM with peepholes
(file-local create_lstm default, :12; layer name `f_lstm_agg`), frame-level head — mirrors `modelzoo/adenet_v2_4.py:32-121`."""
from .. import init
from .adenet_v2_1 import build_raw_diff


def create_model(ae, diff_ae, input_shape, input_var, mask_shape, mask_var, diff_shape, diff_var, lstm_size=250,
                 win=None, output_classes=26, fusiontype='concat', w_init_fn=init.Orthogonal(), use_peepholes=True):
    return build_raw_diff(ae, diff_ae, input_shape, input_var, mask_shape, mask_var, diff_shape, diff_var, lstm_size, win,
                          output_classes, fusiontype, w_init_fn, use_peepholes, 'frame_lstm')
