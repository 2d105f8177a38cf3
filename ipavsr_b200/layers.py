"""Layer graph with Lasagne's names, constructor arguments, traversal order and parameter layout.

This is the host-side mirror of the reference's graph-building API for the AdeNet/DeltaNet hot path: the
builders in `ipavsr_b200/modelzoo/*` compose these objects exactly as the reference composes Lasagne layers
(`modelzoo/adenet_v2.py:30-92` etc.), so layer names, `get_all_layers` order and the `get_all_param_values`
pickle layout stay drop-in.  Layers hold *no arithmetic*: they are topology +
parameters.  The numbers are produced by the sm_100a kernels behind `ipavsr_b200.engine` (the product)
or, in tests only, by the NumPy oracle under `oracle/`.

Custom layers of the reference (`custom/layers.py`): `DeltaLayer` (:105), `AdaptiveElemwiseSumLayer` (:178).
"""
from collections import OrderedDict, deque

import numpy as np

from . import init
from . import nonlinearities


# --------------------------------------------------------------------------------------------------
# Symbolic placeholders (stand-ins for theano.tensor variables; they only carry identity + dtype)
# --------------------------------------------------------------------------------------------------
class Var(object):
    """Opaque placeholder for a `theano.tensor` input variable (`runners/2stream_dct.py:246-250`)."""

    def __init__(self, name=None, ndim=None, dtype='float32'):
        self.name, self.ndim, self.dtype = name, ndim, dtype

    def __repr__(self):
        return 'Var(%s)' % self.name


class Param(object):
    """A named parameter tensor with Lasagne-style tags.

    `value` is the host master copy until an engine binds the parameter to its device arena; afterwards
    `get_value()` reads the device and `set_value()` writes it (see `engine.ParamArena`).  A parameter can be bound by
    several engines (a function compiled on an intermediate layer after training, `create_pretrained_model` reusing
    layers): every arena carries a modification stamp per parameter, `get_value()` reads the most recently modified
    copy, `set_value()` writes all of them, and an engine refreshes its copy from a newer one before it runs.
    """

    def __init__(self, value, name, tags):
        self._host = np.array(value, dtype=np.float32, order='C')
        self.name = name
        self.tags = set(tags)
        self.shape = self._host.shape
        self._bindings = []           # weak references to the arenas that hold a device copy

    def arenas(self):
        live = [r() for r in self._bindings]
        if any(a is None for a in live):
            self._bindings = [r for r, a in zip(self._bindings, live) if a is not None]
        return [a for a in live if a is not None]

    @property
    def _binding(self):
        """(arena, slot) of the most recently modified device copy, or None while the parameter lives on the host."""
        live = self.arenas()
        if not live:
            return None
        return max(live, key=lambda a: a.stamp_of(self)), self

    def get_value(self):
        b = self._binding
        if b is not None:
            return b[0].read(self)
        return self._host.copy()

    def set_value(self, v):
        v = np.asarray(v, dtype=np.float32)
        if v.shape != self.shape:
            raise ValueError('mismatch: parameter %s has shape %r but value has shape %r'
                             % (self.name, self.shape, v.shape))
        self._host = np.array(v, dtype=np.float32, order='C')
        live = self.arenas()
        if live:
            stamp = live[0].next_stamp()
            for arena in live:
                arena.write(self, self._host, stamp=stamp)

    def __repr__(self):
        return self.name


class Layer(object):
    def __init__(self, incoming, name=None):
        if isinstance(incoming, tuple):
            self.input_shape = incoming
            self.input_layer = None
        else:
            self.input_shape = incoming.output_shape
            self.input_layer = incoming
        self.name = name
        self.params = OrderedDict()

    @property
    def output_shape(self):
        return self.get_output_shape_for(self.input_shape)

    def get_output_shape_for(self, input_shape):
        return input_shape

    def add_param(self, spec, shape, name=None, **tags):
        if name is not None and self.name is not None:
            name = '%s.%s' % (self.name, name)
        tags['trainable'] = tags.get('trainable', True)
        tags['regularizable'] = tags.get('regularizable', True)
        if isinstance(spec, Param):
            param = spec
        else:
            if callable(spec):
                value = spec(shape)
            else:
                value = np.asarray(spec, dtype=np.float32)
            value = np.asarray(value, dtype=np.float32)
            if value.shape != tuple(shape):
                raise ValueError('parameter %s: expected shape %r, got %r' % (name, tuple(shape), value.shape))
            param = Param(value, name, [])
        self.params[param] = set(tag for tag, v in tags.items() if v)
        param.tags = self.params[param]
        return param

    def get_params(self, unwrap_shared=True, **tags):
        result = list(self.params.keys())
        only = set(tag for tag, v in tags.items() if v)
        if only:
            result = [p for p in result if not (only - self.params[p])]
        exclude = set(tag for tag, v in tags.items() if not v)
        if exclude:
            result = [p for p in result if not (self.params[p] & exclude)]
        return result


class MergeLayer(Layer):
    def __init__(self, incomings, name=None):
        self.input_shapes = [i if isinstance(i, tuple) else i.output_shape for i in incomings]
        self.input_layers = [None if isinstance(i, tuple) else i for i in incomings]
        self.name = name
        self.params = OrderedDict()

    @property
    def output_shape(self):
        return self.get_output_shape_for(self.input_shapes)


class InputLayer(Layer):
    def __init__(self, shape, input_var=None, name=None):
        self.shape = tuple(shape)
        self.input_var = input_var if input_var is not None else Var(name, len(self.shape))
        self.name = name
        self.params = OrderedDict()

    @property
    def output_shape(self):
        return self.shape


class ReshapeLayer(Layer):
    """`ReshapeLayer(l, (-1, D))` / `(batch, seqlen, F)` / `(-1, seqlen, C)`; symbolic entries are opaque."""

    def __init__(self, incoming, shape, name=None):
        super(ReshapeLayer, self).__init__(incoming, name)
        self.shape = tuple(shape)
        if not isinstance(self.shape[-1], (int, np.integer)):
            raise ValueError('the last (feature) dimension of a ReshapeLayer must be a concrete int')

    def get_output_shape_for(self, input_shape):
        return tuple(None if not isinstance(s, (int, np.integer)) or s == -1 else int(s) for s in self.shape)


class DenseLayer(Layer):
    def __init__(self, incoming, num_units, W=init.GlorotUniform(), b=init.Constant(0.),
                 nonlinearity=nonlinearities.rectify, name=None):
        super(DenseLayer, self).__init__(incoming, name)
        self.nonlinearity = nonlinearities.identity if nonlinearity is None else nonlinearity
        self.num_units = int(num_units)
        num_inputs = int(np.prod(self.input_shape[1:]))
        self.W = self.add_param(W, (num_inputs, self.num_units), name='W')
        self.b = None if b is None else self.add_param(b, (self.num_units,), name='b', regularizable=False)

    def get_output_shape_for(self, input_shape):
        return (input_shape[0], self.num_units)


class BatchNormLayer(Layer):
    """Lasagne BatchNormLayer on a 2-D input, axes=(0,) (`modelzoo/adenet_v1.py:82`)."""

    def __init__(self, incoming, epsilon=1e-4, alpha=0.1, name=None):
        super(BatchNormLayer, self).__init__(incoming, name)
        self.epsilon, self.alpha = epsilon, alpha
        f = (int(self.input_shape[-1]),)
        self.beta = self.add_param(init.Constant(0), f, 'beta', trainable=True, regularizable=False)
        self.gamma = self.add_param(init.Constant(1), f, 'gamma', trainable=True, regularizable=True)
        self.mean = self.add_param(init.Constant(0), f, 'mean', trainable=False, regularizable=False)
        self.inv_std = self.add_param(init.Constant(1), f, 'inv_std', trainable=False, regularizable=False)


class DropoutLayer(Layer):
    def __init__(self, incoming, p=0.5, rescale=True, name=None):
        super(DropoutLayer, self).__init__(incoming, name)
        self.p, self.rescale = p, rescale


class DeltaLayer(Layer):
    """Appends 1st/2nd-order delta coefficients: (N,T,F) -> (N,T,3F)  (`custom/layers.py:105-121`)."""

    def __init__(self, incoming, window, name=None):
        super(DeltaLayer, self).__init__(incoming, name)
        self.window = window

    def get_output_shape_for(self, input_shape):
        return input_shape[0], input_shape[1], input_shape[-1] * 3


class Gate(object):
    def __init__(self, W_in=init.Normal(0.1), W_hid=init.Normal(0.1), W_cell=init.Normal(0.1),
                 b=init.Constant(0.), nonlinearity=nonlinearities.sigmoid):
        self.W_in, self.W_hid, self.b = W_in, W_hid, b
        if W_cell is not None:
            self.W_cell = W_cell
        self.nonlinearity = nonlinearities.identity if nonlinearity is None else nonlinearity


class LSTMLayer(MergeLayer):
    """Lasagne LSTMLayer as the reference configures it.  Note the Lasagne default
    `peepholes=True`: builders that omit the kwarg (`adenet_v1.py:26-42`, `adenet_v3.py:113-143`,
    `lstm_classifier_baseline.py:15-51`) get peepholes."""

    def __init__(self, incoming, num_units, ingate=None, forgetgate=None, cell=None, outgate=None,
                 nonlinearity=nonlinearities.tanh, cell_init=init.Constant(0.), hid_init=init.Constant(0.),
                 backwards=False, learn_init=False, peepholes=True, gradient_steps=-1, grad_clipping=0,
                 unroll_scan=False, precompute_input=True, mask_input=None, only_return_final=False,
                 name=None):
        incomings = [incoming]
        self.mask_incoming_index = -1
        if mask_input is not None:
            incomings.append(mask_input)
            self.mask_incoming_index = 1
        super(LSTMLayer, self).__init__(incomings, name)
        ingate = ingate if ingate is not None else Gate()
        forgetgate = forgetgate if forgetgate is not None else Gate()
        cell = cell if cell is not None else Gate(W_cell=None, nonlinearity=nonlinearities.tanh)
        outgate = outgate if outgate is not None else Gate()
        if gradient_steps != -1 or only_return_final or nonlinearity is not nonlinearities.tanh:
            raise ValueError('LSTMLayer: only full BPTT, tanh output and full-sequence output are supported')
        for g, want in ((ingate, 'sigmoid'), (forgetgate, 'sigmoid'), (cell, 'tanh'), (outgate, 'sigmoid')):
            if g.nonlinearity.name != want:
                raise ValueError('LSTMLayer: gate nonlinearity %s not supported here' % g.nonlinearity.name)
        self.num_units = int(num_units)
        self.backwards = backwards
        self.learn_init = learn_init
        self.peepholes = peepholes
        self.grad_clipping = float(grad_clipping)
        num_inputs = int(np.prod(self.input_shapes[0][2:]))
        self.num_inputs = num_inputs
        H = self.num_units

        def add_gate_params(gate, gate_name):
            return (self.add_param(gate.W_in, (num_inputs, H), name='W_in_to_%s' % gate_name),
                    self.add_param(gate.W_hid, (H, H), name='W_hid_to_%s' % gate_name),
                    self.add_param(gate.b, (H,), name='b_%s' % gate_name, regularizable=False))

        self.W_in_to_ingate, self.W_hid_to_ingate, self.b_ingate = add_gate_params(ingate, 'ingate')
        self.W_in_to_forgetgate, self.W_hid_to_forgetgate, self.b_forgetgate = \
            add_gate_params(forgetgate, 'forgetgate')
        self.W_in_to_cell, self.W_hid_to_cell, self.b_cell = add_gate_params(cell, 'cell')
        self.W_in_to_outgate, self.W_hid_to_outgate, self.b_outgate = add_gate_params(outgate, 'outgate')
        if self.peepholes:
            self.W_cell_to_ingate = self.add_param(ingate.W_cell, (H,), name='W_cell_to_ingate')
            self.W_cell_to_forgetgate = self.add_param(forgetgate.W_cell, (H,), name='W_cell_to_forgetgate')
            self.W_cell_to_outgate = self.add_param(outgate.W_cell, (H,), name='W_cell_to_outgate')
        self.cell_init = self.add_param(cell_init, (1, H), name='cell_init',
                                        trainable=learn_init, regularizable=False)
        self.hid_init = self.add_param(hid_init, (1, H), name='hid_init',
                                       trainable=learn_init, regularizable=False)

    def get_output_shape_for(self, input_shapes):
        s = input_shapes[0]
        return s[0], s[1], self.num_units


class ElemwiseSumLayer(MergeLayer):
    def __init__(self, incomings, coeffs=1, name=None):
        super(ElemwiseSumLayer, self).__init__(incomings, name)
        if coeffs != 1:
            raise ValueError('ElemwiseSumLayer: fixed coeffs are not used by the reference and not supported')

    def get_output_shape_for(self, input_shapes):
        return input_shapes[0]


class AdaptiveElemwiseSumLayer(MergeLayer):
    """sum_s alpha_s * x_s with scalar trainable alpha_s initialised to 1.0 and *always* multiplied
    (`custom/layers.py:217-224`; the `coeff != 1` test there is an identity test on a shared variable).
    The parameters keep the shared variable's own name `adacoeff{i}` (no layer prefix)."""

    def __init__(self, incomings, name=None):
        super(AdaptiveElemwiseSumLayer, self).__init__(incomings, name)
        self.coeffs = []
        for i in range(len(incomings)):
            p = Param(np.float32(1.0).reshape(()), 'adacoeff%d' % i, [])
            self.coeffs.append(self.add_param(p, (), trainable=True, scaling_param=True))

    def get_output_shape_for(self, input_shapes):
        return input_shapes[0]


class ConcatLayer(MergeLayer):
    def __init__(self, incomings, axis=1, name=None):
        super(ConcatLayer, self).__init__(incomings, name)
        self.axis = axis
        rank = len(self.input_shapes[0])
        if axis not in (rank - 1, -1):
            raise ValueError('ConcatLayer: only feature-axis concatenation is supported (axis=%r)' % (axis,))

    def get_output_shape_for(self, input_shapes):
        out = list(input_shapes[0])
        out[-1] = sum(s[-1] for s in input_shapes)
        return tuple(out)


class SliceLayer(Layer):
    """`SliceLayer(l, indices=-1, axis=1)` -> x[:, -1]: the *padded* last time index."""

    def __init__(self, incoming, indices, axis=-1, name=None):
        super(SliceLayer, self).__init__(incoming, name)
        if indices != -1 or axis != 1 or len(self.input_shape) != 3:
            raise ValueError('SliceLayer: only indices=-1, axis=1 on a (N,T,F) input is supported')
        self.indices, self.axis = indices, axis

    def get_output_shape_for(self, input_shape):
        return input_shape[0], input_shape[2]


# --------------------------------------------------------------------------------------------------
# lasagne.layers helper functions
# --------------------------------------------------------------------------------------------------
def get_all_layers(layer, treat_as_input=None):
    """Depth-first post-order over input layers in declaration order, each layer once (Lasagne
    `layers/helper.py:get_all_layers`)."""
    try:
        queue = deque(layer)
    except TypeError:
        queue = deque([layer])
    seen, done, result = set(), set(), []
    if treat_as_input is not None:
        seen.update(treat_as_input)
    while queue:
        l = queue[0]
        if l is None:
            queue.popleft()
        elif l not in seen:
            seen.add(l)
            if hasattr(l, 'input_layers'):
                queue.extendleft(reversed(l.input_layers))
            elif getattr(l, 'input_layer', None) is not None:
                queue.appendleft(l.input_layer)
        else:
            queue.popleft()
            if l not in done:
                result.append(l)
                done.add(l)
    return result


def get_all_params(layer, unwrap_shared=True, **tags):
    result, seen = [], set()
    for l in get_all_layers(layer):
        for p in l.get_params(**tags):
            if p not in seen:
                seen.add(p)
                result.append(p)
    return result


def count_params(layer, **tags):
    return int(sum(int(np.prod(p.shape)) for p in get_all_params(layer, **tags)))


def get_all_param_values(layer, **tags):
    """List of host ndarrays in Lasagne order — the payload `utils/io.py:40-42` pickles."""
    return [p.get_value() for p in get_all_params(layer, **tags)]


def set_all_param_values(layer, values, **tags):
    params = get_all_params(layer, **tags)
    if len(params) != len(values):
        raise ValueError('mismatch: got %d values to set %d parameters' % (len(values), len(params)))
    for p, v in zip(params, values):
        p.set_value(v)


def get_output_shape(layer_or_layers, input_shapes=None):
    if isinstance(layer_or_layers, (list, tuple)):
        return [l.output_shape for l in layer_or_layers]
    return layer_or_layers.output_shape


class OutputExpr(object):
    """What `las.layers.get_output(net, deterministic=...)` returns here: a handle naming the output layer
    and the mode.  `ipavsr_b200.function` turns it into device work."""

    def __init__(self, layer, deterministic):
        self.layer, self.deterministic = layer, bool(deterministic)


def get_output(layer, inputs=None, deterministic=False, **kwargs):
    return OutputExpr(layer, deterministic)
