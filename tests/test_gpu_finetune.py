"""-m gpu parity tests of SURVEY §8f rank 4: auto-encoder fine-tuning (nolearn-style NeuralNet over the encoder GEMMs,
squared-error + L2 objective, Nesterov momentum) against the CPU oracle."""
import numpy as np
import pytest

import gpu_util as G
from oracle import ops
from oracle.net import OracleNet

SIZES = (64, 48, 32, 16, 8, 16, 32, 48, 64)


def _make_net(rng, lr=0.001, mom=0.05, l2c=0.005, sizes=SIZES, **kw):
    from ipavsr_b200 import layers as L
    from ipavsr_b200.nonlinearities import sigmoid, linear
    from ipavsr_b200.custom.nolearn_net import NeuralNet
    from ipavsr_b200.custom.objectives import squared_error
    from ipavsr_b200.custom.updates import nesterov_momentum
    specs = [(L.InputLayer, {'name': 'input', 'shape': (None, sizes[0])})]
    n = len(sizes) - 1
    for i in range(n):
        act = linear if i in (n // 2 - 1, n - 1) else sigmoid          # l4 (bottleneck) and output are linear
        W = (rng.normal(size=(sizes[i], sizes[i + 1])) / np.sqrt(sizes[i])).astype(np.float32)
        b = (0.1 * rng.normal(size=(sizes[i + 1],))).astype(np.float32)
        specs.append((L.DenseLayer, {'name': 'output' if i == n - 1 else 'l%d' % (i + 1), 'num_units': sizes[i + 1],
                                     'nonlinearity': act, 'W': W, 'b': b}))
    return NeuralNet(layers=specs, max_epochs=2, objective_loss_function=squared_error, update=nesterov_momentum,
                     regression=True, update_learning_rate=lr, update_momentum=mom, objective_l2=l2c, **kw)


@pytest.mark.gpu
def test_squared_error_and_l2_kernels_through_the_c_abi():
    import torch
    rng = np.random.default_rng(0)
    for M, F, ld in ((37, 50, 56), (1000, 1200, 1200), (5, 7, 7)):
        p = np.zeros((M, ld), 'float32'); t = np.zeros((M, ld), 'float32')
        p[:, :F] = rng.normal(size=(M, F)); t[:, :F] = rng.normal(size=(M, F))
        dp, dt_ = G.dev(p), G.dev(t)
        dg = G.zeros((M, ld)); loss = G.zeros((4,))
        gs = 1.0 / (M * F)
        G.call('ipavsr_squared_error', dp.data_ptr(), ld, dt_.data_ptr(), ld, loss.data_ptr(), dg.data_ptr(), ld, M, F, gs,
               G.stream())
        want, dwant = ops.squared_error_mean(p[:, :F].astype(np.float64), t[:, :F].astype(np.float64), np.float64)
        assert abs(G.host(loss)[0] * gs - want) <= 1e-6 * want
        got = G.host(dg)
        np.testing.assert_allclose(got[:, :F], dwant, rtol=1e-6, atol=1e-9)
        assert not got[:, F:].any()
    # l2 over a 3-tensor arena: the middle tensor has coefficient 0
    n = 256 * 7 + 100
    w = rng.normal(size=n).astype('float32'); g0 = rng.normal(size=n).astype('float32')
    ids = np.array([0, 0, 0, 1, 1, 2, 2, 2], dtype=np.int32)
    coef = np.array([0.005, 0.0, 0.25], dtype=np.float32)
    dw, dg, loss, dcoef, dids = G.dev(w), G.dev(g0), G.zeros((4,)), G.dev(coef), G.dev(ids)
    G.call('ipavsr_l2_penalty', dw.data_ptr(), dg.data_ptr(), n, dcoef.data_ptr(), dids.data_ptr(), loss.data_ptr(), 3.0,
           G.stream())
    c = coef[ids[np.arange(n) // 256]].astype(np.float64)
    np.testing.assert_allclose(G.host(dg), g0 + 2 * c * w, rtol=1e-6, atol=1e-7)
    assert abs(G.host(loss)[0] - 3.0 * (c * w.astype(np.float64) ** 2).sum()) <= 1e-5 * 3.0 * (c * w ** 2).sum()


@pytest.mark.gpu
def test_autoencoder_objective_and_gradients_match_the_oracle():
    from ipavsr_b200 import layers as L
    rng = np.random.default_rng(1)
    net = _make_net(rng)
    net.initialize()
    X = rng.normal(size=(96, SIZES[0])).astype(np.float32)
    params = L.get_all_params(net._out, trainable=True)
    ref_loss, ref_out, ref_grads = OracleNet(net._out, np.float64).loss_and_grads(
        {'input': X.reshape(96, 1, -1)}, 0, X, None, 'squared_error', deterministic=False, l2=0.005)
    recon = net.predict(X)
    assert recon.shape == X.shape
    assert np.abs(recon - ref_out.reshape(X.shape)).max() <= 1e-4 * np.abs(ref_out).max()
    val = float(net.eval_iter_(X.reshape(96, 1, -1), X))
    loss = float(net.train_iter_(X.reshape(96, 1, -1), X))
    assert abs(loss - float(ref_loss)) <= 1e-4 * abs(float(ref_loss))
    assert abs(val - float(ref_loss)) <= 1e-4 * abs(float(ref_loss))
    grads = net.train_iter_.engine.param_grads(params)
    gmax = max(float(np.abs(r).max()) for r in ref_grads)
    for g, r, p in zip(grads, ref_grads, params):
        assert np.abs(g - r).max() <= 2e-3 * max(float(np.abs(r).max()), 2e-2 * gmax), p.name


@pytest.mark.gpu
def test_fit_follows_the_oracle_training_loop():
    """Two epochs of NeuralNet.fit (unshuffled 128-row batches, first-fold validation split, Nesterov momentum) against the
    float32 oracle loop; then the fine-tuned encoder layers are read back the way the builders do."""
    from ipavsr_b200 import layers as L
    from ipavsr_b200.custom.nolearn_net import train_split
    rng = np.random.default_rng(2)
    net = _make_net(rng, lr=0.01, mom=0.05)
    net.initialize()
    X = rng.normal(size=(700, SIZES[0])).astype(np.float32)
    params = L.get_all_params(net._out, trainable=True)
    w0 = [p.get_value().copy() for p in params]
    # oracle loop on a twin network
    twin = _make_net(np.random.default_rng(2), lr=0.01, mom=0.05)
    twin.initialize()
    tparams = L.get_all_params(twin._out, trainable=True)
    for p, w in zip(tparams, w0):
        p.set_value(w)
    tr, va = train_split(len(X), 0.2)
    assert len(va) == 140 and va[0] == 0 and tr[0] == 140
    state = {'vel': [np.zeros_like(w) for w in w0]}
    hist = []
    for ep in range(2):
        tl, tn = [], []
        for i in range(0, len(tr), 128):
            xb = X[tr][i:i + 128]
            on = OracleNet(twin._out, np.float32)
            loss, _, grads = on.loss_and_grads({'input': xb.reshape(len(xb), 1, -1)}, 0, xb, None, 'squared_error',
                                               deterministic=False, l2=0.005)
            vals = [p.get_value() for p in tparams]
            ops.sgd_momentum_step(vals, grads, state, 0.01, 0.05, nesterov=True)
            for p, v in zip(tparams, vals):
                p.set_value(v)
            tl.append(float(loss)); tn.append(len(xb))
        hist.append(np.average(tl, weights=tn))
    net.fit(X, X)
    assert len(net.train_history_) == 2
    for h, want in zip(net.train_history_, hist):
        assert abs(h['train_loss'] - want) <= 2e-4 * want
        assert np.isfinite(h['valid_loss'])
    assert net.train_history_[1]['train_loss'] < net.train_history_[0]['train_loss']
    for p, tp, w in zip(params, tparams, w0):
        moved = np.abs(tp.get_value() - w).max()
        assert np.abs(p.get_value() - tp.get_value()).max() <= 5e-3 * moved + 1e-6, p.name
    layers = net.get_all_layers()
    assert [l.name for l in layers] == ['input'] + ['l%d' % i for i in range(1, 8)] + ['output']
    assert layers[4].W.get_value().shape == (SIZES[3], SIZES[4])
