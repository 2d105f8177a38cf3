"""Behaviour of the compile-step mirror (`ipavsr_b200.function`) that the reference gets from Theano for free: parameters
shared between compiled functions stay coherent, `updates` honour their parameter list and loss, engine options cannot be
changed silently.  Run with -m gpu on the B200."""
import numpy as np
import pytest

from ipavsr_b200 import layers as L
from ipavsr_b200.engine import Engine, get_engine
from ipavsr_b200.function import function, tensor as T
from ipavsr_b200.custom.objectives import temporal_softmax_loss
from ipavsr_b200.custom import updates as U
from oracle.net import OracleNet
import model_util as MU

pytestmark = pytest.mark.gpu


def _net(seed=3, name='adenet_v2', fusiontype='concat', N=9, T=11):
    rng = np.random.default_rng(seed)
    spec = MU.build(name, rng, C=7, H=12, win=3, fusiontype=fusiontype)
    net = spec['net']
    MU.randomize_params(net, rng)
    xs, mask, lens = MU.make_feed(rng, N, T, spec['dims'])
    y = np.repeat(rng.integers(0, 7, size=(N, 1)), T, 1).astype('int32')
    feed = dict(zip(spec['names'], xs))
    feed['mask'] = mask
    return spec, net, feed, mask, y


def _train_fn(net, spec, upd_fn):
    ins = MU.input_layers(net)
    pred = L.get_output(net, deterministic=False)
    targets = T.imatrix('t')
    cost = temporal_softmax_loss(pred, targets, ins['mask'].input_var)
    order = [ins[n].input_var for n in spec['names']]
    train = function([order[0], targets, ins['mask'].input_var] + order[1:] + [T.iscalar('w')], cost,
                     updates=upd_fn(cost))
    return train, cost, ins


def test_functions_on_sub_networks_see_trained_parameters():
    """Train through the output layer, then compile a function on an intermediate layer (feature extraction, a second
    engine over a subset of the same Params): it must run on the TRAINED weights, get_all_param_values must return them,
    and set_all_param_values must reach the arena the train function uses."""
    spec, net, feed, mask, y = _net()
    params = L.get_all_params(net, trainable=True)
    before = [p.get_value() for p in params]
    train, cost, ins = _train_fn(net, spec, lambda c: U.adam(c, params, learning_rate=1e-2))
    for _ in range(3):
        train(feed['input'], y, mask, feed['dct'], 3)
    after = [p.get_value() for p in params]
    assert all(np.abs(a - b).max() > 0 for a, b in zip(after, before))
    # a second engine on the bottleneck of the same network
    bott = [l for l in L.get_all_layers(net) if l.name == 'bottleneck'][0]
    feat = function([ins['input'].input_var], L.get_output(bott, deterministic=True))
    got = feat(feed['input'])
    # oracle forward of the sub-network with the trained values
    want = OracleNet(bott, np.float64).forward({'input': feed['input']}, 3, deterministic=True)
    assert np.abs(got.reshape(want.shape) - want).max() / np.abs(want).max() < 1e-4
    for a, b in zip(L.get_all_param_values(net), [p.get_value() for p in L.get_all_params(net)]):
        np.testing.assert_array_equal(a, b)
    for p, a in zip(params, after):
        np.testing.assert_array_equal(p.get_value(), a)
    # more training is seen by the second engine on its next call
    train(feed['input'], y, mask, feed['dct'], 3)
    got2 = feat(feed['input'])
    assert np.abs(got2 - got).max() > 0
    want2 = OracleNet(bott, np.float64).forward({'input': feed['input']}, 3, deterministic=True)
    assert np.abs(got2.reshape(want2.shape) - want2).max() / np.abs(want2).max() < 1e-4
    # set_all_param_values (restoring the best epoch, runners/2stream_dct.py:382-383) reaches the training arena
    L.set_all_param_values(net, [v * 0 for v in L.get_all_param_values(net)])
    assert np.abs(feat(feed['input'])).max() == 0
    run, out = train.engine.forward({ins[k]: v for k, v in feed.items()}, 3, deterministic=True)
    probs = train.engine.read(out)
    assert np.allclose(probs, 1.0 / 7, atol=1e-6)


def test_update_parameter_subset_freezes_the_rest():
    spec, net, feed, mask, y = _net(seed=4)
    params = L.get_all_params(net, trainable=True)
    sub = [p for p in params if p.name.split('.')[0] in ('softmax', 'f_lstm_agg', 'b_lstm_agg')]
    assert 0 < len(sub) < len(params)
    before = [p.get_value() for p in params]
    for rule in (lambda c: U.adam(c, sub, learning_rate=1e-2), ):
        train, cost, ins = _train_fn(net, spec, rule)
        train(feed['input'], y, mask, feed['dct'], 3)
    after = [p.get_value() for p in params]
    for p, a, b in zip(params, before, after):
        assert (np.abs(a - b).max() > 0) == (p in sub), p.name


def test_update_subset_must_not_split_a_device_tensor():
    spec, net, feed, mask, y = _net(seed=5)
    params = L.get_all_params(net, trainable=True)
    sub = [p for p in params if p.name == 'f_lstm_agg.W_in_to_ingate']
    train, cost, ins = _train_fn(net, spec, lambda c: U.sgd(c, sub, learning_rate=1e-2))
    with pytest.raises(ValueError):
        train(feed['input'], y, mask, feed['dct'], 3)


def test_updates_built_from_another_loss_are_refused():
    spec, net, feed, mask, y = _net(seed=6)
    params = L.get_all_params(net, trainable=True)
    ins = MU.input_layers(net)
    targets = T.imatrix('t')
    cost_a = temporal_softmax_loss(L.get_output(net, deterministic=False), targets, ins['mask'].input_var)
    cost_b = temporal_softmax_loss(L.get_output(net, deterministic=True), targets, ins['mask'].input_var)
    order = [ins[n].input_var for n in spec['names']]
    args = [order[0], targets, ins['mask'].input_var] + order[1:] + [T.iscalar('w')]
    with pytest.raises(ValueError):
        function(args, cost_a, updates=U.adam(cost_b, params))
    # an equal expression built twice is fine
    cost_c = temporal_softmax_loss(L.get_output(net, deterministic=False), targets, ins['mask'].input_var)
    function(args, cost_a, updates=U.adam(cost_c, params))


def test_engine_options_cannot_change_silently():
    spec, net, feed, mask, y = _net(seed=7)
    eng = get_engine(net, gemm_mode='fp32')
    assert get_engine(net) is eng and get_engine(net, gemm_mode='fp32') is eng
    with pytest.raises(ValueError):
        get_engine(net, gemm_mode='f16x3')
    with pytest.raises(ValueError):
        get_engine(net, packed='force')


def test_default_mode_is_the_tensor_core_mode():
    spec, net, feed, mask, y = _net(seed=8)
    import os
    if 'IPAVSR_GEMM_MODE' not in os.environ:
        assert Engine(net).gemm_mode == 4        # f16x3: the benchmarked arithmetic is the shipped default


@pytest.mark.parametrize('name,fusiontype', [('adenet_v2', 'concat'), ('adenet_v1', 'sum'), ('deltanet', 'sum')])
def test_small_batch_cuda_graph_step_matches_eager(name, fusiontype, monkeypatch):
    """Launch-bound small batches (the reference's own batch sizes) replay ONE CUDA graph per step; losses and parameters
    after several steps on changing batches equal the eager path's."""
    from ipavsr_b200.custom.objectives import categorical_crossentropy

    def run(graph):
        monkeypatch.setenv('IPAVSR_GRAPH', '1' if graph else '0')
        spec, net, feed, mask, y = _net(seed=12, name=name, fusiontype=fusiontype, N=10, T=12)
        level = spec['level']
        ins = MU.input_layers(net)
        pred = L.get_output(net, deterministic=False)
        params = L.get_all_params(net, trainable=True)
        if level == 'frame':
            tg = T.imatrix('t')
            cost = temporal_softmax_loss(pred, tg, ins['mask'].input_var)
        else:
            tg = T.ivector('t')
            cost = T.mean(categorical_crossentropy(pred, tg))
        order = [ins[n].input_var for n in spec['names']]
        train = function([order[0], tg, ins['mask'].input_var] + order[1:] + [T.iscalar('w')], cost,
                         updates=U.adam(cost, params, learning_rate=1e-2))
        assert (train.engine.graph_mode == 'auto') == graph
        rng = np.random.default_rng(99)
        losses = []
        for step in range(4):
            xs, m2, _ = MU.make_feed(rng, 10, 12, spec['dims'])          # a new batch (new lengths) every step
            yy = rng.integers(0, 7, size=10).astype('int32')
            yy = yy if level == 'seq' else np.repeat(yy[:, None], 12, 1).astype('int32')
            losses.append(float(train(xs[0], yy, m2, *xs[1:], 3)))
        if graph:
            assert len(train.engine._graphs) == 1 and not train.engine._graph_failed
        return losses, [p.get_value() for p in params]

    l_g, p_g = run(True)
    l_e, p_e = run(False)
    np.testing.assert_allclose(l_g, l_e, rtol=2e-5)
    for a, b in zip(p_g, p_e):
        assert np.abs(a - b).max() <= 2e-4 * max(1.0, np.abs(b).max())


@pytest.mark.parametrize('name,fusiontype', [('adenet_v2', 'concat'), ('adenet_v2', 'sum'), ('adenet_3stream', 'concat'),
                                             ('adenet_4stream', 'concat')])
@pytest.mark.parametrize('graph', [False, True])
def test_branch_streams_match_single_stream(name, fusiontype, graph, monkeypatch):
    """Small batches run every input branch (encoder -> DeltaLayer -> LSTM) on a stream of its own, forward and backward
    (Engine._branch_enter); losses and parameters after several steps equal the single-stream issue order.  The kernels
    and their arguments are the same, so the only admissible difference is the order of the float atomics of k-split
    weight gradients."""
    def run(branches):
        monkeypatch.setenv('IPAVSR_GRAPH', '1' if graph else '0')
        monkeypatch.setenv('IPAVSR_BRANCH_STREAMS', '1' if branches else '0')
        spec, net, feed, mask, y = _net(seed=21, name=name, fusiontype=fusiontype, N=10, T=12)
        ins = MU.input_layers(net)
        pred = L.get_output(net, deterministic=False)
        params = L.get_all_params(net, trainable=True)
        tg = T.imatrix('t') if spec['level'] == 'frame' else T.ivector('t')
        if spec['level'] == 'frame':
            cost = temporal_softmax_loss(pred, tg, ins['mask'].input_var)
        else:
            from ipavsr_b200.custom.objectives import categorical_crossentropy
            cost = T.mean(categorical_crossentropy(pred, tg))
        order = [ins[n].input_var for n in spec['names']]
        train = function([order[0], tg, ins['mask'].input_var] + order[1:] + [T.iscalar('w')], cost,
                         updates=U.adam(cost, params, learning_rate=1e-2))
        eng = train.engine
        assert eng.branch_mode == branches
        assert eng._n_branches == len(spec['names'])
        rng = np.random.default_rng(5)
        losses = []
        for step in range(4):
            xs, m2, _ = MU.make_feed(rng, 10, 12, spec['dims'])
            yy = rng.integers(0, 7, size=10).astype('int32')
            yy = yy if spec['level'] == 'seq' else np.repeat(yy[:, None], 12, 1).astype('int32')
            losses.append(float(train(xs[0], yy, m2, *xs[1:], 3)))
        if branches:
            assert len(eng._branch_streams) == eng._n_branches          # every branch really got its stream
        return losses, [p.get_value() for p in params]

    l_b, p_b = run(True)
    l_s, p_s = run(False)
    np.testing.assert_allclose(l_b, l_s, rtol=2e-5)
    for a, b in zip(p_b, p_s):
        assert np.abs(a - b).max() <= 2e-4 * max(1.0, np.abs(b).max())
