"""2-stream late-fusion AdeNet — mirrors `modelzoo/adenet_2stream.py`: `create_model` (:116-208) and
`create_pretrained_model` (:12-114, sub-stream LSTMs loaded from LSTM `.mat` dicts, optionally forward+backward summed)."""
from .. import init
from . import _nstream


def create_pretrained_model(s1_ae, s1_lstm, s2_ae, s2_lstm, s1_shape, s1_var, s2_shape, s2_var, mask_shape, mask_var,
                            lstm_size=250, win=None, output_classes=26, fusiontype='concat',
                            w_init_fn=init.Orthogonal(), use_peepholes=True, use_blstm_substream=False):
    return _nstream.build([s1_ae, s2_ae], [s1_shape, s2_shape], [s1_var, s2_var], mask_shape, mask_var, lstm_size, win,
                          output_classes, fusiontype, w_init_fn, use_peepholes, lstm_weights=[s1_lstm, s2_lstm],
                          use_blstm_substream=use_blstm_substream)


def create_model(s1_ae, s2_ae, s1_shape, s1_var, s2_shape, s2_var, mask_shape, mask_var, lstm_size=250, win=None,
                 output_classes=26, fusiontype='concat', w_init_fn=init.Orthogonal(), use_peepholes=True):
    return _nstream.build([s1_ae, s2_ae], [s1_shape, s2_shape], [s1_var, s2_var], mask_shape, mask_var, lstm_size, win,
                          output_classes, fusiontype, w_init_fn, use_peepholes)
