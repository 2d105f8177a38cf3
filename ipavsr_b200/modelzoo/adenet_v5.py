"""AdeNet v5: the v3 wiring with a `use_adascale` switch (adaptive sum or plain sum) — mirrors
`modelzoo/adenet_v5.py:64-184`; returns (l_out, l_sum1)."""
from . import adenet_v3

create_pretrained_encoder = adenet_v3.create_pretrained_encoder
extract_weights = adenet_v3.extract_weights


def create_model(ae, diff_ae, input_shape, input_var, mask_shape, mask_var, dct_shape, dct_var, diff_shape, diff_var,
                 lstm_size=250, win=None, output_classes=26, use_adascale=False):
    return adenet_v3._build(ae, diff_ae, input_shape, input_var, mask_shape, mask_var, dct_shape, dct_var, diff_shape,
                            diff_var, lstm_size, win, output_classes, 'adasum' if use_adascale else 'sum')
