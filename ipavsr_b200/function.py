"""The compile step of the reference runners, mirrored: `theano.function(inputs, outputs, updates=...)`
(`runners/2stream_dct.py:268-279`) becomes `ipavsr_b200.function(inputs, outputs, updates=...)`, returning a
callable with the same positional arguments that runs the B200 engine.

    predictions = layers.get_output(network, deterministic=False)
    cost = temporal_softmax_loss(predictions, targets, mask)
    updates = adam(cost, all_params, learning_rate=lr)
    train = function([inputs1, targets, mask, inputs2, window], cost, updates=updates)
    val_fn = function([inputs1, mask, inputs2, window], test_predictions)

`tensor` provides the placeholder constructors the runners use (`T.tensor3`, `T.matrix`, `T.imatrix`,
`T.ivector`, `T.iscalar`).
"""
import os

import numpy as np

from . import layers as L
from .engine import get_engine, _Nvtx


class _TensorNamespace(object):
    @staticmethod
    def tensor3(name=None, dtype='float32'):
        return L.Var(name, 3, dtype)

    @staticmethod
    def matrix(name=None, dtype='float32'):
        return L.Var(name, 2, dtype)

    @staticmethod
    def imatrix(name=None):
        return L.Var(name, 2, 'int32')

    @staticmethod
    def ivector(name=None):
        return L.Var(name, 1, 'int32')

    @staticmethod
    def iscalar(name=None):
        return L.Var(name, 0, 'int32')

    @staticmethod
    def mean(x):
        if isinstance(x, LossExpr) and x.kind == 'categorical_crossentropy_elemwise':
            return LossExpr('categorical_crossentropy', x.pred, x.targets, None)
        if isinstance(x, LossExpr) and x.kind == 'squared_error_elemwise':
            return LossExpr('squared_error', x.pred, x.targets, None)
        raise TypeError('T.mean is only defined on categorical_crossentropy(...) / squared_error(...) here')


tensor = _TensorNamespace()


class LossExpr(object):
    def __init__(self, kind, pred, targets, mask, l2=0.0):
        self.kind, self.pred, self.targets, self.mask, self.l2 = kind, pred, targets, mask, float(l2)

    def __add__(self, other):
        """`loss + coefficient * regularize_network_params(net, l2)` (nolearn's objective, `avletters/trimodal.py:87`)."""
        from .regularization import Penalty
        if not isinstance(other, Penalty):
            return NotImplemented
        if other.layer is not self.pred.layer:
            raise ValueError('the penalty must be taken over the network the loss is computed on')
        return LossExpr(self.kind, self.pred, self.targets, self.mask, self.l2 + other.coefficient)

    __radd__ = __add__


class UpdateSpec(object):
    def __init__(self, kind, loss, params, learning_rate=None, lr_map=None, **hp):
        if not isinstance(loss, LossExpr):
            raise TypeError('updates need a loss expression')
        self.kind, self.loss, self.params, self.lr, self.lr_map, self.hp = kind, loss, params, learning_rate, lr_map, hp


def _lr_value(lr):
    return float(lr.get_value()) if hasattr(lr, 'get_value') else float(lr)


class shared(object):
    """Minimal stand-in for a theano shared scalar (learning rates that decay per epoch,
    `avletters/trimodal.py:436-437`)."""

    def __init__(self, value, name=None):
        self.value, self.name = np.float32(value), name

    def get_value(self):
        return self.value

    def set_value(self, v):
        self.value = np.float32(v)


def function(inputs, outputs=None, updates=None, allow_input_downcast=True, on_unused_input='raise', **engine_kw):
    if isinstance(outputs, (list, tuple)):
        if len(outputs) != 1:
            raise ValueError('one output expression per function is supported')
        outputs = outputs[0]
    if isinstance(outputs, LossExpr):
        pred = outputs.pred
    elif isinstance(outputs, L.OutputExpr):
        pred = outputs
    else:
        raise TypeError('outputs must come from layers.get_output(...) or a loss built on it')
    eng = get_engine(pred.layer, **engine_kw)
    by_var = {l.input_var: l for l in eng.input_layers}
    slots = []          # per positional argument: ('input', layer) | ('targets',) | ('mask-only',) | ('window',)
    for v in inputs:
        if v in by_var:
            slots.append(('input', by_var[v]))
        elif isinstance(outputs, LossExpr) and v is outputs.targets:
            slots.append(('targets', None))
        elif isinstance(v, L.Var) and v.ndim == 0:
            slots.append(('window', None))
        else:
            raise ValueError('input %r is not used by the network (unused inputs are an error, as in Theano)' % (v,))
    need = [l for l in eng.input_layers]
    have = [s[1] for s in slots if s[0] == 'input']
    for l in need:
        if l not in have:
            raise ValueError('missing input for InputLayer %r' % (l.name,))
    mask_layer = None
    if isinstance(outputs, LossExpr) and outputs.mask is not None:
        mask_layer = by_var.get(outputs.mask)
        if mask_layer is None:
            raise ValueError('the loss mask must be the network mask input')
    is_loss = isinstance(outputs, LossExpr)
    train = updates is not None
    loss_name = {'temporal_softmax': 'temporal_softmax', 'categorical_crossentropy': 'categorical_crossentropy',
                 'squared_error': 'squared_error'}.get(outputs.kind) if is_loss else None
    if is_loss and loss_name is None:
        raise TypeError('wrap categorical_crossentropy(...) / squared_error(...) in T.mean(...)')
    l2 = outputs.l2 if is_loss else 0.0
    subset = None
    if train:
        # Theano derives the updates from THEIR loss and THEIR parameter list (`lasagne.updates.adam(cost, params, lr)`):
        # the gradients here come from `outputs`, so the two must be the same expression, and a parameter list that is
        # not the full trainable set becomes a per-tensor learning-rate table (absent tensors: rate 0, as in adam_vlr)
        ul = updates.loss
        same = ul is outputs or (is_loss and ul.kind == outputs.kind and ul.pred.layer is outputs.pred.layer and
                                 ul.pred.deterministic == outputs.pred.deterministic and ul.targets is outputs.targets and
                                 ul.mask is outputs.mask and ul.l2 == outputs.l2)
        if not same:
            raise ValueError('the updates were built from a different loss expression than the function output: the step '
                             'would silently ignore it (pass the same cost to both)')
        if updates.lr_map is None and updates.params is not None:
            every = L.get_all_params(pred.layer, trainable=True)
            unknown = [p for p in updates.params if p not in every]
            if unknown:
                raise ValueError('parameters %r are not trainable parameters of this network' % (unknown,))
            if set(updates.params) != set(every):
                subset = list(updates.params)

    def fn(*args, **kw):
        if len(args) != len(slots):
            raise TypeError('expected %d arguments, got %d' % (len(slots), len(args)))
        feed, window, y = {}, None, None
        for (kind, layer), a in zip(slots, args):
            if kind == 'input':
                feed[layer] = a
            elif kind == 'targets':
                y = a
            else:
                window = int(a)
        dropout_masks = kw.get('dropout_masks')
        if not is_loss:
            run, out = eng.forward(feed, window, pred.deterministic, train=False, dropout_masks=dropout_masks)
            lay = pred.layer
            if kw.get('device_output'):
                # the result stays in HBM as a torch tensor (on-device evaluation, utils/evaluate.py)
                res = eng.read_device(out)
                return res.reshape(run.N, run.T, -1) if len(lay.output_shape) == 3 else res
            _run_deferred()
            res = eng.read(out)
            if len(lay.output_shape) == 3:
                res = res.reshape(run.N, run.T, -1)
            return res
        mask = feed[mask_layer] if mask_layer is not None else None
        if not train:
            run, out = eng.forward(feed, window, pred.deterministic, train=False, dropout_masks=dropout_masks)
            return eng.loss_only(out, loss_name, y, mask, l2=l2)
        graphed = False
        graph_ok = not dropout_masks and eng.graph_eligible(feed, y, pred.deterministic, l2)
        # Early loss read-back (Engine._loss_early_copy): not with an L2 term (added to the loss after the backward pass) and
        # not for a graph replay.  A deferred prefetch is then staged FIRST: this call returns as soon as its loss is final,
        # the host has ~3 ms of slack per step, and an upload issued before the step's kernels is the one order in which the
        # end-to-end step stays at the device-resident time (7.08 ms; staged after them it settles at 8.0 ms, and with the
        # late read an up-front staging delays the enqueue: 8.6 ms — tools/host_time.py)
        # (Not when the deferred upload is large — three padded host streams, 204 MB, are link-bound for longer than a step:
        # staged first they cost 10.0 ms per step, staged last with the early return 7.8 .. 9.6 ms erratically, staged last
        # with the late read a steady 8.7 ms, which is what such calls keep.)
        big = _pending_bytes() > (128 << 20)
        eng._early_loss_ok = not l2 and not graph_ok and not big
        if eng._early_loss_ok and eng.early_loss:
            _run_deferred()
        if graph_ok:
            with _Nvtx('forward + loss + backward (CUDA graph)'):
                graphed = eng.graph_step(feed, window, y, mask_layer, loss_name, pred.deterministic)
        if not graphed:
            with _Nvtx('forward'):
                run, out = eng.forward(feed, window, pred.deterministic, train=True, dropout_masks=dropout_masks, targets=y)
            eng._ar_enabled = not l2          # an L2 penalty is added to the finished gradient arena: all-reduce afterwards
            with _Nvtx('loss + backward'):
                eng.loss_and_backward(run, out, loss_name, y, run.vals[mask_layer] if mask_layer is not None else None,
                                      count=float(np.asarray(mask).sum())
                                      if (mask is not None and not hasattr(mask, 'is_cuda')) else None)
                if l2:
                    eng.l2_penalty(l2)
        with _Nvtx('gradient all-reduce'):
            eng.allreduce_grads()
        u = updates
        lr = _lr_value(u.lr) if u.lr is not None else 0.0
        lr_map = u.lr_map if subset is None else {p: lr for p in subset}
        with _Nvtx('update (%s)' % u.kind):
            eng.optim_step(u.kind, lr, params=u.params, lr_map=lr_map, **u.hp)
        _run_deferred()
        return eng.read_loss()

    pending = []

    def _pending_bytes():
        """Host bytes the deferred prefetches will upload (device tensors and derived streams cost nothing)."""
        n = 0
        for feed in pending:
            for a in feed.values():
                if hasattr(a, 'is_cuda'):
                    n += 0 if a.is_cuda else a.numel() * a.element_size()
                elif hasattr(a, 'nbytes'):
                    n += int(a.nbytes)
        return n

    def _run_deferred():
        while pending:
            eng.prefetch(pending.pop(0))

    def prefetch(*args, **kw):
        """Stage the host inputs of a LATER call (same positional arguments as the call itself) on the engine's copy stream:
        the copy overlaps whatever the device is doing, and the matching call picks the staged buffers up instead of
        copying again.  `defer=True` postpones the staging to the next call of this function, which performs it where it
        costs the step nothing (before its own kernels when the call returns early on its loss — the usual case —, behind
        them when it reads the loss after the update) — the recommended double-buffered loop:

            train.prefetch(*batch[0])
            for i in range(n):
                if i + 1 < n: train.prefetch(*batch[i + 1], defer=True)
                cost = train(*batch[i])

        The host buffers must stay unchanged until the matching call has returned (pinned memory is read asynchronously).
        """
        if len(args) != len(slots):
            raise TypeError('expected %d arguments, got %d' % (len(slots), len(args)))
        feed = {layer: a for (kind, layer), a in zip(slots, args) if kind == 'input'}
        if kw.get('defer'):
            # staged by the NEXT call of this function (see fn: first thing with the early loss read-back, otherwise after it
            # has enqueued its own kernels and before it waits for its result)
            pending.append(feed)
        else:
            eng.prefetch(feed)

    fn.engine = eng
    fn.prefetch = prefetch
    return fn
