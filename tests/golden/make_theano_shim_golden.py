"""Generates tests/golden/theano_shim.npz by EXECUTING the reference's own Theano-level sources —

    utils/signal.py        append_delta_coeff  (the DeltaLayer's arithmetic, row a2)
    custom/objectives.py   temporal_softmax_loss                          (row a8)
    custom/updates.py      adam_vlr, generate_lr_map                      (row a9)

— imported unmodified from /root/reference (read-only) with `theano`, `theano.tensor` and `lasagne` replaced by the small
NumPy shim below.  Theano itself cannot be installed here, so this is not the reference's compiled graph; it
is the reference's source code run statement by statement on NumPy arrays, with the dtype rules that matter made explicit:

  * int32 * float32 -> float64 (Theano's upcast, NumPy's promotion): the per-theta term of `delta_theta` is float64, the
    accumulator is rounded back to float32 every theta (`utils/signal.py:19-21`) — the quirk the CUDA kernel reproduces
    bit for bit;
  * Python float literals stay float32 next to float32 operands (Theano's autocast, NumPy's weak scalars);
  * `T.constant(1)` is int8 in Theano so that `one - beta2**t` and `one - beta1` stay float32 (`custom/updates.py:79-80`);
    the shim's constant is the float32 value 1 (NumPy would promote int8 next to a Python float to float64).

`theano.scan` is an eager loop; `theano.shared` hands out persistent boxes in creation order, so calling `adam_vlr` once per
step re-binds to the state of the previous step exactly as the compiled update dictionary does.

Run in the build container only:   python tests/golden/make_theano_shim_golden.py
"""
import importlib.util
import os
import sys
import types

import numpy as np

REF = '/root/reference'


# ---------------------------------------------------------------------------------------------------------------------
# the shim
# ---------------------------------------------------------------------------------------------------------------------
class Shared(object):
    """theano.shared: a named mutable box that behaves like its value in arithmetic."""

    def __init__(self, value, name=None, broadcastable=None):
        self.value = np.array(value)
        self.name = name
        self.broadcastable = broadcastable if broadcastable is not None else (False,) * self.value.ndim

    def get_value(self, borrow=False):
        return self.value

    def set_value(self, v):
        self.value = np.array(v, dtype=self.value.dtype)

    @property
    def shape(self):
        return self.value.shape

    def _v(self, o):
        return o.value if isinstance(o, Shared) else o

    def __add__(self, o): return self.value + self._v(o)
    def __radd__(self, o): return self._v(o) + self.value
    def __sub__(self, o): return self.value - self._v(o)
    def __rsub__(self, o): return self._v(o) - self.value
    def __mul__(self, o): return self.value * self._v(o)
    def __rmul__(self, o): return self._v(o) * self.value
    def __truediv__(self, o): return self.value / self._v(o)
    def __rtruediv__(self, o): return self._v(o) / self.value
    def __pow__(self, o): return self.value ** self._v(o)
    def __rpow__(self, o): return self._v(o) ** self.value
    def __hash__(self): return id(self)
    def __eq__(self, o): return self is o


class SharedFactory(object):
    """First pass: creates boxes.  `replay()`: the next pass gets the SAME boxes back in creation order (what a compiled
    Theano function does implicitly: its update dictionary refers to the shared variables made at graph-build time)."""

    def __init__(self):
        self.made, self.cursor = [], None

    def __call__(self, value, name=None, broadcastable=None, **kw):
        if self.cursor is not None:
            box = self.made[self.cursor]
            self.cursor += 1
            return box
        box = Shared(value, name, broadcastable)
        self.made.append(box)
        return box

    def replay(self):
        self.cursor = 0


def scan(fn, sequences=None, outputs_info=None, non_sequences=None, **kw):
    seqs = sequences if isinstance(sequences, (list, tuple)) else [sequences]
    non = list(non_sequences) if isinstance(non_sequences, (list, tuple)) else ([] if non_sequences is None else [non_sequences])
    prev, outs = outputs_info, []
    for i in range(len(seqs[0])):
        args = [s[i] for s in seqs] + ([prev] if outputs_info is not None else []) + non
        r = fn(*args)
        outs.append(np.array(r, copy=True))
        if outputs_info is not None:
            prev = r
    return np.stack(outs), {}


def install(shared_factory):
    theano = types.ModuleType('theano')
    tt = types.ModuleType('theano.tensor')
    extra = types.ModuleType('theano.tensor.extra_ops')
    extra.repeat = lambda x, n: np.repeat(x, n)
    tt.extra_ops = extra
    tt.arange = lambda a, b=None, dtype='int64': np.arange(a, b, dtype=dtype) if b is not None else np.arange(a, dtype=dtype)
    tt.zeros_like = np.zeros_like
    tt.concatenate = lambda xs, axis=0: np.concatenate(xs, axis=axis)
    tt.sum = lambda x, axis=None, keepdims=False: np.sum(x, axis=axis, keepdims=keepdims)
    tt.max = lambda x, axis=None, keepdims=False: np.max(x, axis=axis, keepdims=keepdims)
    tt.exp, tt.log = np.exp, np.log
    tt.sqrt = lambda x: np.sqrt(x.value if isinstance(x, Shared) else x)
    # T.constant(1) is an int8 TensorConstant in Theano: next to a Python float (autocast to floatX) it gives float32.  An
    # np.int8 next to a Python float gives float64 in NumPy, so the shim hands out the value 1 as float32 — same arithmetic
    tt.constant = lambda v: np.float32(v)
    tt.add = np.add
    theano.tensor = tt
    theano.scan = scan
    theano.shared = shared_factory
    lasagne = types.ModuleType('lasagne')
    lutils = types.ModuleType('lasagne.utils')
    lutils.floatX = lambda v: np.float32(v)
    lutils.unroll_scan = None
    lupd = types.ModuleType('lasagne.updates')
    lupd.get_or_compute_grads = lambda loss_or_grads, params: list(loss_or_grads)     # gradients are passed in
    lasagne.utils, lasagne.updates = lutils, lupd
    for name, mod in (('theano', theano), ('theano.tensor', tt), ('theano.tensor.extra_ops', extra), ('lasagne', lasagne),
                      ('lasagne.utils', lutils), ('lasagne.updates', lupd)):
        sys.modules[name] = mod


def load(relpath, modname):
    spec = importlib.util.spec_from_file_location(modname, os.path.join(REF, relpath))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def main():
    factory = SharedFactory()
    install(factory)
    signal = load('utils/signal.py', 'ref_signal')
    objectives = load('custom/objectives.py', 'ref_objectives')
    updates = load('custom/updates.py', 'ref_updates')
    rng = np.random.default_rng(20261018)
    out = {}

    # ---- a2: DeltaLayer = theano.scan(append_delta_coeff, sequences=input, non_sequences=window)  (custom/layers.py:114-117)
    cases = []
    for k, (N, T, F, theta) in enumerate([(3, 11, 5, 1), (2, 16, 7, 2), (2, 23, 4, 4), (2, 40, 6, 9), (1, 5, 3, 9), (2, 9, 50, 3)]):
        x = (rng.normal(size=(N, T, F)) * rng.choice([1e-3, 1.0, 37.0])).astype(np.float32)
        res, _ = scan(signal.append_delta_coeff, sequences=x, non_sequences=theta)
        assert res.shape == (N, T, 3 * F), res.shape
        out['delta_x_%d' % k], out['delta_y_%d' % k], out['delta_theta_%d' % k] = x, res.astype(np.float32), np.int32(theta)
        assert res.dtype == np.float32
        cases.append(k)
    out['delta_cases'] = np.array(cases, np.int32)

    # ---- a8: temporal_softmax_loss(x, y, mask): x are the network's (already soft-maxed) outputs (trimodal.py:327)
    for k, (N, T, V) in enumerate([(4, 9, 26), (7, 40, 10), (1, 3, 2)]):
        logits = rng.normal(size=(N, T, V)) * 2.0
        probs = np.exp(logits - logits.max(2, keepdims=True))
        probs = (probs / probs.sum(2, keepdims=True)).astype(np.float32)
        y = rng.integers(0, V, size=(N, T)).astype(np.int32)
        lens = rng.integers(1, T + 1, size=N)
        mask = (np.arange(T)[None, :] < lens[:, None]).astype(np.uint8)
        loss = objectives.temporal_softmax_loss(probs.copy(), y, mask)
        out['tsl_probs_%d' % k], out['tsl_y_%d' % k], out['tsl_mask_%d' % k] = probs, y, mask
        out['tsl_loss_%d' % k] = np.float64(loss)
    out['tsl_cases'] = np.int32(3)

    # ---- a9: generate_lr_map + adam_vlr over 4 steps with per-layer rates (runners: custom/updates.py:10-32, :35-99)
    names = ['fc1.W', 'fc1.b', 'lstm_s1.W_in_to_ingate', 'output.W']
    shapes = [(6, 5), (5,), (5, 8), (8, 3)]
    params = [Shared((rng.normal(size=s) * 0.3).astype(np.float32), name=n) for n, s in zip(names, shapes)]
    lr_map = updates.generate_lr_map(params, {'fc1': 1e-4, 'lstm_s1': 2e-3}, 1e-3)
    assert [lr_map[p] for p in params] == [1e-4, 1e-4, 2e-3, 1e-3]
    out['adam_lrs'] = np.array([lr_map[p] for p in params], np.float64)
    for i, p in enumerate(params):
        out['adam_p0_%d' % i] = p.get_value().copy()
    steps = 4
    for s in range(steps):
        grads = [(rng.normal(size=sh) * (10.0 ** rng.integers(-4, 1))).astype(np.float32) for sh in shapes]
        if s:
            factory.replay()                                   # the same t_prev / m_prev / v_prev as the step before
        upd = updates.adam_vlr(grads, params, lr_map)
        new = [(var, np.array(val)) for var, val in upd.items()]
        for var, val in new:                                   # Theano applies all updates at once, from the OLD values
            assert np.asarray(val).dtype == np.float32, (var.name, np.asarray(val).dtype)
            var.set_value(val)
        for i, (g, p) in enumerate(zip(grads, params)):
            out['adam_g_%d_%d' % (s, i)] = g
            out['adam_p_%d_%d' % (s, i)] = p.get_value().copy()
    out['adam_steps'], out['adam_nparams'] = np.int32(steps), np.int32(len(params))
    t_prev = factory.made[0]
    assert float(t_prev.get_value()) == float(steps)

    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'theano_shim.npz')
    np.savez_compressed(path, **out)
    print('wrote', path, '(%d arrays)' % len(out))


if __name__ == '__main__':
    main()
