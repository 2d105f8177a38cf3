"""End-to-end parity of every builder on the GPU against the NumPy oracle: probabilities, argmax, loss, every
parameter gradient, and a few optimiser steps (SURVEY §8d parity gates).  Run with -m gpu on the B200."""
import numpy as np
import pytest

from ipavsr_b200 import layers as L
from ipavsr_b200.engine import Engine
from ipavsr_b200.function import function, tensor as T
from ipavsr_b200.custom.objectives import temporal_softmax_loss, categorical_crossentropy
from ipavsr_b200.custom import updates as U
from oracle.net import OracleNet
from oracle import ops
import model_util as MU

pytestmark = pytest.mark.gpu


def _seed(name):
    """Deterministic per-builder seed (str hashes are randomised per process).  It matters: about 1 seed in 75 of a rectify
    network puts one unit's pre-activation within float32 rounding of 0, where the float32 path and the float64 oracle take
    different branches and one weight-gradient column moves by percents."""
    import zlib
    return zlib.crc32(name.encode()) % 1000


def _case(name, seed, fusiontype='sum', N=9, T=11, H=12, C=7, win=3):
    rng = np.random.default_rng(seed)
    spec = MU.build(name, rng, C=C, H=H, win=win, fusiontype=fusiontype)
    net = spec['net']
    MU.randomize_params(net, rng)
    xs, mask, lens = MU.make_feed(rng, N, T, spec['dims'])
    y1 = rng.integers(0, C, size=N).astype('int32')
    y = y1 if spec['level'] == 'seq' else np.repeat(y1[:, None], T, 1).astype('int32')
    feed_names = dict(zip(spec['names'], xs))
    feed_names['mask'] = mask
    dm = MU.dropout_masks_for(net, rng, N, T)
    return spec, net, feed_names, mask, y, dm, win


def _oracle(net, feed, win, y, mask, level, dm, dt=np.float64):
    o = OracleNet(net, dt)
    loss = 'categorical_crossentropy' if level == 'seq' else 'temporal_softmax'
    return o.loss_and_grads(feed, win, y, mask, loss, deterministic=False, dropout_masks=dm, update_bn=False)


def _fusions(name):
    return ['sum', 'adasum', 'concat'] if name in ('adenet_v2', 'adenet_v3', 'adenet_3stream', 'adenet_4stream') else ['sum']


# gradient gate of the small-network sweeps: GRAD_TOL of the per-tensor max (analytically-zero gradients against the scale
# of the others).  SURVEY 8(d) asks for 1e-4; the measured worst case over all builders, fusion types and both modes is
# printed by the tests (-s) and recorded in profiles/.
GRAD_TOL = 1e-4


@pytest.mark.parametrize('mode', ['fp32', 'f16x3'])
@pytest.mark.parametrize('name', MU.ALL)
def test_forward_backward_parity(name, mode):
    """Every builder in the CUDA-core cross-check mode and in the shipped default (f16x3 + tensor-core LSTM; at these toy
    sizes the GEMMs below the tensor-core threshold run the exact FP32 kernel, the recurrences run on the tensor cores)."""
    _check_forward_backward(name, _fusions(name), mode)


def _check_forward_backward(name, fusiontypes, mode='fp32'):
    for fusiontype in fusiontypes:
        spec, net, feed, mask, y, dm, win = _case(name, _seed(name), fusiontype)
        eng = Engine(net, gemm_mode=mode)
        ins = MU.input_layers(net)
        dfeed = {ins[k]: v for k, v in feed.items()}
        run, out = eng.forward(dfeed, win, deterministic=False, train=True, dropout_masks=dm, update_bn=False)
        # the oracle's backward takes the device's branch at rectify units within rounding of 0 (counted, and required to
        # be at rounding distance: model_util.rectify_aligner)
        align, flips = MU.rectify_aligner(net, run, mask.shape[0], mask.shape[1])
        loss_name = 'categorical_crossentropy' if spec['level'] == 'seq' else 'temporal_softmax'
        loss_ref, out_ref, grads_ref = OracleNet(net, np.float64).loss_and_grads(
            feed, win, y, mask, loss_name, deterministic=False, dropout_masks=dm, update_bn=False, after_forward=align)
        probs = eng.read(out).reshape(out_ref.shape)
        rel = np.abs(probs - out_ref).max() / np.abs(out_ref).max()
        assert rel < 1e-4, (name, fusiontype, 'probs', rel)
        assert (probs.argmax(-1) == out_ref.argmax(-1)).all()
        eng.loss_and_backward(run, out, loss_name, y, mask, count=float(mask.sum()))
        loss = eng.read_loss()
        assert abs(loss - loss_ref) < 1e-4 * abs(loss_ref), (name, loss, loss_ref)
        params = L.get_all_params(net, trainable=True)
        errs = MU.grad_errors(params, eng.param_grads(params), grads_ref)
        for p, err in errs:
            assert err < GRAD_TOL, (name, fusiontype, mode, p.name, err, flips)
        print('PARITY %s %s %s: probs rel %.2e, worst grad err %.2e, rectify flips %s'
              % (name, fusiontype, mode, rel, max(e for _, e in errs), flips))


@pytest.mark.parametrize('mode', ['f16x3', 'tf32x3'])
def test_tensor_core_modes_at_reference_sizes(mode):
    """The fp32-parity tensor-core GEMM modes on the real layer sizes (D=1200, DBNF 2000-1000-500-50, DCT 90, LSTM-250,
    26 classes, T=40): same gates as the CUDA-core fp32 mode — probabilities 1e-4, identical argmax, gradients 1e-4."""
    from ipavsr_b200 import modelzoo, init
    rng = np.random.default_rng(77)
    np.random.seed(77)
    D, Dd, H, C, win, N, T_ = 1200, 90, 250, 26, 9, 24, 40
    ae = MU.ae_tuple(rng, D, shapes=(2000, 1000, 500, 50), acts=('sigmoid', 'sigmoid', 'sigmoid', 'linear'))
    net, _ = modelzoo.adenet_v2.create_model(ae, (None, None, D), T.tensor3('x'), (None, None),
                                             T.matrix('mask', dtype='uint8'), (None, None, Dd), T.tensor3('dct'), H, win, C,
                                             'concat', init.Orthogonal(), True)
    MU.randomize_params(net, rng)
    lens = rng.integers(12, T_ + 1, size=N)
    lens[0] = T_
    xs, mask, _ = MU.make_feed(rng, N, T_, [D, Dd], lens=lens)
    y = np.repeat(rng.integers(0, C, size=(N, 1)), T_, 1).astype('int32')
    feed = {'input': xs[0], 'dct': xs[1], 'mask': mask}
    loss_ref, out_ref, grads_ref = OracleNet(net, np.float64).loss_and_grads(feed, win, y, mask, 'temporal_softmax',
                                                                            deterministic=False, update_bn=False)
    eng = Engine(net, gemm_mode=mode)
    ins = MU.input_layers(net)
    run, out = eng.forward({ins[k]: v for k, v in feed.items()}, win, deterministic=False, train=True, update_bn=False)
    probs = eng.read(out).reshape(out_ref.shape)
    rel = np.abs(probs - out_ref).max() / np.abs(out_ref).max()
    assert rel < 1e-4, (mode, 'probs', rel)
    assert (probs.argmax(-1) == out_ref.argmax(-1)).all()
    eng.loss_and_backward(run, out, 'temporal_softmax', y, mask, count=float(mask.sum()))
    assert abs(eng.read_loss() - loss_ref) < 1e-4 * abs(loss_ref)
    params = L.get_all_params(net, trainable=True)
    errs = MU.grad_errors(params, eng.param_grads(params), grads_ref)
    for p, err in errs:
        assert err < GRAD_TOL, (mode, p.name, err)
    print('PARITY adenet_v2 concat N=24 reference sizes, mode %s: probs rel %.2e, worst grad err %.2e'
          % (mode, rel, max(e for _, e in errs)))


def test_deterministic_eval_and_val_fn_api():
    """The runners' call convention: val_fn(X, mask, X2, window) -> (N,T,C) probabilities."""
    spec, net, feed, mask, y, dm, win = _case('adenet_v2', 5, 'concat')
    o = OracleNet(net, np.float64)
    ref = o.forward(feed, win, deterministic=True)
    ins = MU.input_layers(net)
    window = T.iscalar('theta')
    val_fn = function([ins['input'].input_var, ins['mask'].input_var, ins['dct'].input_var, window],
                      L.get_output(net, deterministic=True))
    got = val_fn(feed['input'], mask, feed['dct'], win)
    assert got.shape == ref.shape and got.dtype == np.float32
    assert np.abs(got - ref).max() / np.abs(ref).max() < 1e-4
    assert (got.argmax(-1) == ref.argmax(-1)).all()
    # double-buffered input staging: a prefetched call returns the same result and consumes the staged buffers
    val_fn.prefetch(feed['input'], mask, feed['dct'], win)
    assert len(val_fn.engine._prefetched) == 1
    got_pf = val_fn(feed['input'], mask, feed['dct'], win)
    np.testing.assert_array_equal(got_pf, got)
    assert len(val_fn.engine._prefetched) == 0
    # float64 / wide-int inputs are downcast like allow_input_downcast=True
    got2 = val_fn(feed['input'].astype('float64'), mask.astype('int64'), feed['dct'].astype('float64'), win)
    np.testing.assert_allclose(got2, got, rtol=1e-6, atol=1e-7)


@pytest.mark.parametrize('name,rule', [('adenet_v2', 'adam'), ('adenet_v3', 'adadelta'), ('deltanet', 'nesterov'),
                                       ('adenet_v1', 'adam')])
def test_training_steps_match_oracle(name, rule):
    spec, net, feed, mask, y, dm, win = _case(name, 11, 'adasum' if name == 'adenet_v2' else 'sum')
    params = L.get_all_params(net, trainable=True)
    host = [p.get_value() for p in params]
    level = spec['level']
    ins = MU.input_layers(net)
    pred = L.get_output(net, deterministic=False)
    targets = T.imatrix('t') if level == 'frame' else T.ivector('t')
    if level == 'frame':
        cost = temporal_softmax_loss(pred, targets, ins['mask'].input_var)
    else:
        cost = T.mean(categorical_crossentropy(pred, targets))
    if rule == 'adam':
        upd = U.adam(cost, params, learning_rate=1e-2)
    elif rule == 'adadelta':
        upd = U.adadelta(cost, params, learning_rate=1.0)
    else:
        upd = U.nesterov_momentum(cost, params, learning_rate=1e-2, momentum=0.9)
    order = [ins[n].input_var for n in spec['names']]
    window = T.iscalar('theta')
    train = function([order[0], targets, ins['mask'].input_var] + order[1:] + [window], cost, updates=upd)
    # oracle trajectory on host copies (BN running stats follow along through set_value in the oracle)
    import copy
    losses_ref, losses = [], []
    st = {'t': np.float32(0), 'm': [np.zeros_like(h) for h in host], 'v': [np.zeros_like(h) for h in host],
          'acc': [np.zeros_like(h) for h in host], 'dacc': [np.zeros_like(h) for h in host],
          'vel': [np.zeros_like(h) for h in host]}
    # reference run uses a second, identical network object so device state is untouched
    spec2, net2, *_ = _case(name, 11, 'adasum' if name == 'adenet_v2' else 'sum')
    params2 = L.get_all_params(net2, trainable=True)
    for p2, h in zip(params2, host):
        p2.set_value(h)
    for step in range(3):
        lr, _, g = OracleNet(net2, np.float64).loss_and_grads(
            feed, win, y, mask, 'categorical_crossentropy' if level == 'seq' else 'temporal_softmax',
            deterministic=False, dropout_masks=dm, update_bn=True)
        losses_ref.append(float(lr))
        g = [gi.astype('float32') for gi in g]
        cur = [p2.get_value() for p2 in params2]
        if rule == 'adam':
            ops.adam_step(cur, g, st, [1e-2] * len(cur))
        elif rule == 'adadelta':
            ops.adadelta_step(cur, g, st, 1.0)
        else:
            ops.sgd_momentum_step(cur, g, st, 1e-2, 0.9, True)
        for p2, c in zip(params2, cur):
            p2.set_value(c)
        losses.append(float(train(feed[spec['names'][0]], y, mask, *[feed[n] for n in spec['names'][1:]], win,
                                  dropout_masks=dm)))
    np.testing.assert_allclose(losses, losses_ref, rtol=2e-3)
    gmax = max(np.abs(gi).max() for gi in g)
    for p, p2, gi in zip(params, params2, g):
        if np.abs(gi).max() < 1e-4 * gmax:
            continue    # analytically-zero gradient (bias in front of BatchNorm): Adam amplifies pure rounding noise
        a, b = p.get_value(), p2.get_value()
        assert np.abs(a - b).max() < 2e-3 * max(1.0, np.abs(b).max()), p.name


def test_adam_vlr_and_param_roundtrip():
    spec, net, feed, mask, y, dm, win = _case('deltanet_majority_vote', 3)
    params = L.get_all_params(net, trainable=True)
    before = [p.get_value() for p in params]
    ins = MU.input_layers(net)
    pred = L.get_output(net, deterministic=False)
    targets = T.imatrix('t')
    cost = temporal_softmax_loss(pred, targets, ins['mask'].input_var)
    lr_map = U.generate_lr_map(params, {'fc1': 0.0, 'fc2': 0.0, 'fc3': 0.0, 'bottleneck': 0.0}, 1e-2)
    train = function([ins['input'].input_var, targets, ins['mask'].input_var, T.iscalar('w')], cost,
                     updates=U.adam_vlr(cost, params, lr_map))
    train(feed['input'], y, mask, win)
    after = [p.get_value() for p in params]
    for p, a, b in zip(params, before, after):
        frozen = p.name.split('.')[0] in ('fc1', 'fc2', 'fc3', 'bottleneck')
        assert (np.abs(a - b).max() == 0) == frozen, p.name
    # get/set_all_param_values round trip through the device arena (the pickle layout)
    vals = L.get_all_param_values(net)
    L.set_all_param_values(net, [v * 0 + 1 for v in vals])
    assert all((v == 1).all() for v in L.get_all_param_values(net))
    L.set_all_param_values(net, vals)
    for a, b in zip(vals, L.get_all_param_values(net)):
        np.testing.assert_array_equal(a, b)
