"""LSTM factory helpers and custom layers — mirrors `custom/layers.py:10-80,105-121,178-228`."""
import numpy as np

from ..layers import LSTMLayer, Gate, DeltaLayer, AdaptiveElemwiseSumLayer   # noqa: F401 (re-exported)


def create_lstm(l_incoming, l_mask, hidden_units, cell_parameters, gate_parameters, name, use_peepholes=False):
    """`custom/layers.py:10-25`."""
    if cell_parameters is None:
        cell_parameters = Gate()
    if gate_parameters is None:
        gate_parameters = Gate()
    return LSTMLayer(l_incoming, hidden_units, peepholes=use_peepholes, mask_input=l_mask,
                     ingate=gate_parameters, forgetgate=gate_parameters, cell=cell_parameters,
                     outgate=gate_parameters, learn_init=True, grad_clipping=5., name=name)


def create_pretrained_lstm(lstm_weights, prefix, l_incoming, l_mask, hidden_units, cell_parameters,
                           gate_parameters, name, use_peepholes=False, backwards=False):
    """`custom/layers.py:28-52`: build an LSTM and load the 12 `{prefix}_w_*`/`{prefix}_b_*` arrays of an
    LSTM `.mat` (peepholes and learned inits are not carried, as in the reference)."""
    l_lstm = LSTMLayer(l_incoming, hidden_units, peepholes=use_peepholes, mask_input=l_mask,
                       ingate=gate_parameters, forgetgate=gate_parameters, cell=cell_parameters,
                       outgate=gate_parameters, learn_init=True, grad_clipping=5., name=name,
                       backwards=backwards)
    for g in ('cell', 'forgetgate', 'ingate', 'outgate'):
        getattr(l_lstm, 'W_hid_to_' + g).set_value(
            np.asarray(lstm_weights['{}_w_hid_to_{}'.format(prefix, g)]).astype('float32'))
        getattr(l_lstm, 'W_in_to_' + g).set_value(
            np.asarray(lstm_weights['{}_w_in_to_{}'.format(prefix, g)]).astype('float32'))
        getattr(l_lstm, 'b_' + g).set_value(
            np.asarray(lstm_weights['{}_b_{}'.format(prefix, g)]).astype('float32').reshape((-1,)))
    return l_lstm


def create_blstm(l_incoming, l_mask, hidden_units, cell_parameters, gate_parameters, name, use_peepholes=False):
    """`custom/layers.py:55-80`: forward `f_<name>` and backward `b_<name>` LSTMs over the same input."""
    if cell_parameters is None:
        cell_parameters = Gate()
    if gate_parameters is None:
        gate_parameters = Gate()
    l_lstm = LSTMLayer(l_incoming, hidden_units, peepholes=use_peepholes, mask_input=l_mask,
                       ingate=gate_parameters, forgetgate=gate_parameters, cell=cell_parameters,
                       outgate=gate_parameters, learn_init=True, grad_clipping=5., name='f_{}'.format(name))
    l_lstm_back = LSTMLayer(l_incoming, hidden_units, ingate=gate_parameters, peepholes=use_peepholes,
                            mask_input=l_mask, forgetgate=gate_parameters, cell=cell_parameters,
                            outgate=gate_parameters, learn_init=True, grad_clipping=5., backwards=True,
                            name='b_{}'.format(name))
    return l_lstm, l_lstm_back
