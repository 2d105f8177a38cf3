"""CPU ORACLE — test infrastructure only (see oracle/ops.py header for the parity status).

Interprets a layer graph built by the `modelzoo` builders (plain Python topology objects from
`ipavsr_b200/layers.py`; no arithmetic is borrowed from the product) with the NumPy ops of `oracle/ops.py`:
forward, loss, full backward to every parameter, and the update rules.  This plays the role of
`theano.function(...)` + `T.grad` in the reference runners (`runners/2stream_dct.py:263-279`).
"""
import numpy as np

from . import ops


def _lstm_params(layer, dt):
    g = lambda p: p.get_value().astype(dt)
    p = {'W_in': np.concatenate([g(layer.W_in_to_ingate), g(layer.W_in_to_forgetgate),
                                 g(layer.W_in_to_cell), g(layer.W_in_to_outgate)], axis=1),
         'W_hid': np.concatenate([g(layer.W_hid_to_ingate), g(layer.W_hid_to_forgetgate),
                                  g(layer.W_hid_to_cell), g(layer.W_hid_to_outgate)], axis=1),
         'b': np.concatenate([g(layer.b_ingate), g(layer.b_forgetgate), g(layer.b_cell), g(layer.b_outgate)]),
         'cell_init': g(layer.cell_init).reshape(-1), 'hid_init': g(layer.hid_init).reshape(-1)}
    if layer.peepholes:
        p['peep'] = np.stack([g(layer.W_cell_to_ingate), g(layer.W_cell_to_forgetgate),
                              g(layer.W_cell_to_outgate)])
    return p


class OracleNet(object):
    """forward/backward over the Lasagne-ordered layer list of `output_layer`."""

    def __init__(self, output_layer, dt=np.float32):
        from ipavsr_b200 import layers as L          # topology classes only
        self.L = L
        self.out = output_layer
        self.layers = L.get_all_layers(output_layer)
        self.dt = dt

    # inputs: dict InputLayer-name -> ndarray ; dropout_masks: dict layer-name -> {0,1} array
    def forward(self, inputs, window, deterministic=True, dropout_masks=None, update_bn=True):
        L, dt = self.L, self.dt
        vals, caches = {}, {}
        self.window = int(window)
        for l in self.layers:
            if isinstance(l, L.InputLayer):
                v = np.asarray(inputs[l.name])
                vals[l] = v if v.dtype == np.uint8 else v.astype(dt)
            elif isinstance(l, L.ReshapeLayer):
                x = vals[l.input_layer]
                F = l.shape[-1]
                if len(l.shape) == 2:
                    vals[l] = x.reshape(-1, F)
                else:
                    vals[l] = x.reshape(self._N, -1, F)
            elif isinstance(l, L.DenseLayer):
                x = vals[l.input_layer]
                y, c = ops.dense_fwd(x, l.W.get_value(), None if l.b is None else l.b.get_value(),
                                     ops.act_code(l.nonlinearity), dt)
                vals[l], caches[l] = y, c
            elif isinstance(l, L.BatchNormLayer):
                y, c, new = ops.bn_fwd(vals[l.input_layer], l.beta.get_value(), l.gamma.get_value(),
                                       l.mean.get_value(), l.inv_std.get_value(), deterministic,
                                       l.epsilon, l.alpha, dt)
                vals[l], caches[l] = y, c
                if not deterministic and update_bn:
                    l.mean.set_value(new[0].astype(np.float32))
                    l.inv_std.set_value(new[1].astype(np.float32))
            elif isinstance(l, L.DropoutLayer):
                x = vals[l.input_layer]
                if deterministic:
                    vals[l] = x
                else:
                    m = np.asarray(dropout_masks[l.name]).astype(dt).reshape(x.shape)
                    scale = dt(1.0 / (1.0 - l.p)) if l.rescale else dt(1)
                    vals[l], caches[l] = x * m * scale, m * scale
            elif isinstance(l, L.DeltaLayer):
                x = vals[l.input_layer]
                if dt == np.float32:
                    vals[l] = ops.delta_fwd(x, self.window)
                else:
                    D = ops.delta_matrix(x.shape[1], self.window)
                    d = np.einsum('ts,nsf->ntf', D, x)
                    a = np.einsum('ts,nsf->ntf', D, d)
                    vals[l] = np.concatenate([x, d, a], 2)
            elif isinstance(l, L.LSTMLayer):
                x = vals[l.input_layers[0]]
                if l.mask_incoming_index > 0:
                    mask = vals[l.input_layers[1]]
                else:
                    mask = np.ones(x.shape[:2], np.uint8)
                y, c = ops.lstm_fwd(x, mask, _lstm_params(l, dt), l.backwards, dt)
                vals[l], caches[l] = y, c
            elif isinstance(l, L.AdaptiveElemwiseSumLayer):
                xs = [vals[i] for i in l.input_layers]
                cs = [dt(c.get_value()) for c in l.coeffs]
                vals[l] = sum(c * x for c, x in zip(cs, xs))
            elif isinstance(l, L.ElemwiseSumLayer):
                vals[l] = sum(vals[i] for i in l.input_layers)
            elif isinstance(l, L.ConcatLayer):
                vals[l] = np.concatenate([vals[i] for i in l.input_layers], axis=-1)
            elif isinstance(l, L.SliceLayer):
                vals[l] = vals[l.input_layer][:, -1]
            else:
                raise TypeError('oracle: unsupported layer %r' % (l,))
            if isinstance(l, L.InputLayer) and vals[l].ndim == 3:
                self._N = vals[l].shape[0]
        self.vals, self.caches = vals, caches
        return vals[self.out]

    def backward(self, dout):
        """Returns {Param: grad ndarray (Lasagne shape)} for every parameter touched."""
        L, dt = self.L, self.dt
        grads = {}
        gv = {self.out: np.asarray(dout, dt)}

        def acc(layer, g):
            if layer is None or isinstance(layer, L.InputLayer):
                return
            gv[layer] = g if layer not in gv else gv[layer] + g

        for l in reversed(self.layers):
            if l not in gv or isinstance(l, L.InputLayer):
                continue
            g = gv[l]
            if isinstance(l, L.ReshapeLayer):
                acc(l.input_layer, g.reshape(self.vals[l.input_layer].shape))
            elif isinstance(l, L.DenseLayer):
                need_dx = not isinstance(l.input_layer, L.InputLayer)
                dx, dW, db = ops.dense_bwd(g, self.caches[l], l.W.get_value(), ops.act_code(l.nonlinearity),
                                           dt, need_dx=True)
                grads[l.W] = dW
                if l.b is not None:
                    grads[l.b] = db
                acc(l.input_layer, dx)
            elif isinstance(l, L.BatchNormLayer):
                dx, dbeta, dgamma = ops.bn_bwd(g, self.caches[l], l.gamma.get_value(), dt)
                grads[l.beta], grads[l.gamma] = dbeta, dgamma
                acc(l.input_layer, dx)
            elif isinstance(l, L.DropoutLayer):
                acc(l.input_layer, g * self.caches[l] if l in self.caches else g)
            elif isinstance(l, L.DeltaLayer):
                acc(l.input_layer, ops.delta_bwd(g, self.window, dt))
            elif isinstance(l, L.LSTMLayer):
                dx, gr = ops.lstm_bwd(g, self.caches[l], l.grad_clipping, dt)
                H = l.num_units
                for k, name in enumerate(('ingate', 'forgetgate', 'cell', 'outgate')):
                    grads[getattr(l, 'W_in_to_' + name)] = gr['W_in'][:, k * H:(k + 1) * H]
                    grads[getattr(l, 'W_hid_to_' + name)] = gr['W_hid'][:, k * H:(k + 1) * H]
                    grads[getattr(l, 'b_' + name)] = gr['b'][k * H:(k + 1) * H]
                if l.peepholes:
                    grads[l.W_cell_to_ingate] = gr['peep'][0]
                    grads[l.W_cell_to_forgetgate] = gr['peep'][1]
                    grads[l.W_cell_to_outgate] = gr['peep'][2]
                grads[l.cell_init] = gr['cell_init'].reshape(1, H)
                grads[l.hid_init] = gr['hid_init'].reshape(1, H)
                acc(l.input_layers[0], dx)
            elif isinstance(l, L.AdaptiveElemwiseSumLayer):
                for c, i in zip(l.coeffs, l.input_layers):
                    grads[c] = np.asarray((g * self.vals[i]).sum(), dt).reshape(())
                    acc(i, g * dt(c.get_value()))
            elif isinstance(l, L.ElemwiseSumLayer):
                for i in l.input_layers:
                    acc(i, g)
            elif isinstance(l, L.ConcatLayer):
                o = 0
                for i in l.input_layers:
                    w = self.vals[i].shape[-1]
                    acc(i, g[..., o:o + w])
                    o += w
            elif isinstance(l, L.SliceLayer):
                full = np.zeros(self.vals[l.input_layer].shape, dt)
                full[:, -1] = g
                acc(l.input_layer, full)
            else:
                raise TypeError('oracle: unsupported layer %r' % (l,))
        return grads

    # convenience: loss + grads in get_all_params(trainable=True) order
    def loss_and_grads(self, inputs, window, y, mask, loss='temporal_softmax', deterministic=False,
                       dropout_masks=None, update_bn=True, l2=0.0, after_forward=None):
        """`after_forward(self)` (tests only) may edit the forward caches before the backward pass — used to make the
        oracle take the same rectify branch as a float32 run at units whose pre-activation is within rounding of 0."""
        out = self.forward(inputs, window, deterministic, dropout_masks, update_bn)
        if after_forward is not None:
            after_forward(self)
        if loss == 'temporal_softmax':
            val, dout = ops.temporal_softmax_loss(out, y, mask, self.dt)
        elif loss == 'categorical_crossentropy':
            val, dout = ops.categorical_crossentropy_mean(out, y, self.dt)
        elif loss == 'squared_error':
            val, dout = ops.squared_error_mean(out, y, self.dt)
        else:
            raise ValueError(loss)
        grads = self.backward(dout)
        params = self.L.get_all_params(self.out, trainable=True)
        if l2:
            # + l2 * lasagne.regularization.regularize_network_params(net, l2): sum of squares of the regularizable
            # parameters (nolearn objective_l2, avletters/trimodal.py:87)
            for p in self.L.get_all_params(self.out, regularizable=True):
                w = np.asarray(p.get_value(), self.dt)
                val = self.dt(val + self.dt(l2) * (w * w).sum())
                if p in grads:
                    grads[p] = grads[p] + 2 * self.dt(l2) * w.reshape(np.asarray(grads[p]).shape)
        return val, out, [np.asarray(grads[p], self.dt).reshape(p.shape) for p in params]
