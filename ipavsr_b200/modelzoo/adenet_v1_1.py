"""AdeNet v1.1: v1 with dropout before each BLSTM and a 2*lstm_size first BLSTM; returns the output layer only — mirrors
`modelzoo/adenet_v1_1.py:47-104`."""
from . import adenet_v1


def create_model(dbn, input_shape, input_var, mask_shape, mask_var, dct_shape, dct_var, lstm_size=250, win=None,
                 output_classes=26):
    return adenet_v1._build(dbn, input_shape, input_var, mask_shape, mask_var, dct_shape, dct_var, lstm_size, win,
                            output_classes, True)[0]
