"""BLSTM classifier on precomputed features, sequence-level head — mirrors
`modelzoo/lstm_classifier_baseline.py:15-80`.  Its private `create_blstm` omits the `peepholes` kwarg, so
Lasagne's default (`peepholes=True`) applies."""
from .. import init
from ..layers import InputLayer, DenseLayer, SliceLayer, ElemwiseSumLayer, LSTMLayer, Gate
from ..nonlinearities import softmax
from ._common import gates


def create_blstm(l_incoming, l_mask, hidden_units, cell_parameters, gate_parameters, name):
    if cell_parameters is None:
        cell_parameters = Gate()
    if gate_parameters is None:
        gate_parameters = Gate()
    l_lstm = LSTMLayer(l_incoming, hidden_units, mask_input=l_mask, ingate=gate_parameters,
                       forgetgate=gate_parameters, cell=cell_parameters, outgate=gate_parameters,
                       learn_init=True, grad_clipping=5., name='f_{}'.format(name))
    l_lstm_back = LSTMLayer(l_incoming, hidden_units, ingate=gate_parameters, mask_input=l_mask,
                            forgetgate=gate_parameters, cell=cell_parameters, outgate=gate_parameters,
                            learn_init=True, grad_clipping=5., backwards=True, name='b_{}'.format(name))
    return l_lstm, l_lstm_back


def create_model(input_shape, input_var, mask_shape, mask_var, lstm_size=250, output_classes=26,
                 w_init=init.Orthogonal()):
    gate_parameters, cell_parameters = gates(w_init)
    l_in = InputLayer(input_shape, input_var, 'input')
    l_mask = InputLayer(mask_shape, mask_var, 'mask')
    f_lstm, b_lstm = create_blstm(l_in, l_mask, lstm_size, cell_parameters, gate_parameters, 'lstm')
    l_sum = ElemwiseSumLayer([f_lstm, b_lstm], name='sum')
    l_forward_slice1 = SliceLayer(l_sum, -1, 1, name='slice1')
    l_out = DenseLayer(l_forward_slice1, num_units=output_classes, nonlinearity=softmax, name='output')
    return l_out
