"""Issue-order analysis of the device engine (ipavsr_b200/schedule.py): which layers form an input branch (and run on that
branch's CUDA stream), which LSTMs are siblings, and that the backward order stays a reverse topological order.
Pure host logic — runs without a GPU."""
import numpy as np
import pytest

from ipavsr_b200 import layers as L, schedule
import model_util as MU


def _analyse(name, fusiontype='concat'):
    rng = np.random.default_rng(3)
    spec = MU.build(name, rng, C=7, H=12, win=3, fusiontype=fusiontype)
    layers = L.get_all_layers(spec['net'])
    masks = {l.input_layers[1] for l in layers if isinstance(l, L.LSTMLayer) and l.mask_incoming_index > 0}
    return spec, layers, masks, schedule.branch_assignment(layers, masks)


def _ins(l):
    return [i for i in (getattr(l, 'input_layers', None) or [getattr(l, 'input_layer', None)]) if i is not None]


@pytest.mark.parametrize('name,fusion,streams', [('adenet_v2', 'concat', 2), ('adenet_v2', 'sum', 2), ('adenet_v3', 'sum', 3),
                                                 ('adenet_3stream', 'concat', 3), ('adenet_4stream', 'concat', 4)])
def test_one_branch_per_input_stream(name, fusion, streams):
    spec, layers, masks, (branch_of, n, trunk_fed) = _analyse(name, fusion)
    assert n == streams == len(spec['names'])
    inputs = [l for l in layers if isinstance(l, L.InputLayer) and l not in masks]
    assert sorted(branch_of[l] for l in inputs) == list(range(streams))
    for l in layers:
        if l in masks:
            assert branch_of[l] is None                      # the mask is shared: trunk
            continue
        srcs = {branch_of[i] for i in _ins(l) if i not in masks}
        if isinstance(l, L.InputLayer):
            continue
        if srcs == {branch_of[l]} and branch_of[l] is not None:
            continue                                         # a branch layer reads only its own branch (and the mask)
        assert branch_of[l] is None, l.name                  # anything that mixes branches, or follows the trunk, is trunk
    # the network output is behind the fusion; every stream's LSTM hands over to the trunk
    assert branch_of[layers[-1]] is None
    per_branch_lstm = {branch_of[l] for l in layers if isinstance(l, L.LSTMLayer) and branch_of[l] is not None}
    assert per_branch_lstm == set(range(streams))
    assert {branch_of[l] for l in trunk_fed} == set(range(streams))
    for l in trunk_fed:
        assert any(branch_of[c] is None for c in layers if l in _ins(c))


@pytest.mark.parametrize('name', ['deltanet', 'lstm_classifier_baseline'])
def test_single_stream_networks_have_one_branch(name):
    _, layers, masks, (branch_of, n, trunk_fed) = _analyse(name, 'sum')
    assert n == 1                                            # the engine only forks for >= 2 branches
    assert all(b in (0, None) for b in branch_of.values())


@pytest.mark.parametrize('name,fusion', [('adenet_v2', 'concat'), ('adenet_3stream', 'concat'), ('deltanet', 'sum'),
                                         ('adenet_v1', 'sum'), ('adenet_v3', 'sum')])
def test_sibling_lstms_and_backward_order(name, fusion):
    _, layers, masks, (branch_of, n, _) = _analyse(name, fusion)
    sib, order = schedule.lstm_sibling_groups(layers, branch_of)
    assert sorted(map(id, order)) == sorted(map(id, layers))
    pos = {id(l): k for k, l in enumerate(order)}
    for l in layers:                                         # reverse topological: every consumer before its producers
        for i in _ins(l):
            assert pos[id(l)] < pos[id(i)], (l.name, i.name)
    lstms = [l for l in layers if isinstance(l, L.LSTMLayer)]
    for l, others in sib.items():
        for o in others:
            assert o.input_layers[0] is l.input_layers[0] and branch_of[o] == branch_of[l]
            assert l not in _ins(o) and o not in _ins(l)
    # a BLSTM (a forward and a backward LSTM over the same input) is a sibling pair, visited forward-direction first
    pairs = [(a, b) for a in lstms for b in lstms if a is not b and a.input_layers[0] is b.input_layers[0]
             and layers.index(b) == layers.index(a) + 1]
    for a, b in pairs:
        assert sib[a] == (b,) and sib[b] == (a,)
        assert pos[id(a)] < pos[id(b)]
    if name in ('adenet_v2', 'adenet_3stream', 'deltanet'):
        assert pairs, 'these builders end in a BLSTM'
    # switched off: plain reversed order, no groups
    sib0, order0 = schedule.lstm_sibling_groups(layers, branch_of, enabled=False)
    assert not sib0 and order0 == list(reversed(layers))


def test_random_layer_graphs_keep_the_invariants():
    """Random multi-stream graphs (hypothesis): a layer is a branch layer iff it depends on exactly one non-mask input; the
    backward order is a reverse topological order of the whole graph; siblings read the same input and never each other."""
    from hypothesis import given, settings, strategies as hs

    @settings(max_examples=40, deadline=None)
    @given(hs.integers(1, 4), hs.lists(hs.integers(0, 3), min_size=1, max_size=4), hs.booleans(), hs.integers(0, 2 ** 16))
    def run(n_streams, depths, blstm, seed):
        rng = np.random.default_rng(seed)
        mask = L.InputLayer((None, None), name='mask')
        tails = []
        for s in range(n_streams):
            l = L.InputLayer((None, None, 6 + s), name='in%d' % s)
            for d in range(depths[s % len(depths)]):
                l = L.DenseLayer(L.ReshapeLayer(l, (-1, l.output_shape[-1])), 5, name='fc%d_%d' % (s, d))
                l = L.ReshapeLayer(l, (-1, 7, 5))
            if rng.integers(0, 2):
                l = L.DeltaLayer(l, 2, name='delta%d' % s)
            tails.append(L.LSTMLayer(l, 4, mask_input=mask, name='lstm%d' % s))
        agg = tails[0] if len(tails) == 1 else (L.ConcatLayer(tails, axis=2) if rng.integers(0, 2) else L.ElemwiseSumLayer(tails))
        if blstm:
            f = L.LSTMLayer(agg, 3, mask_input=mask, name='f')
            b = L.LSTMLayer(agg, 3, mask_input=mask, backwards=True, name='b')
            agg = L.ElemwiseSumLayer([f, b])
        out = L.DenseLayer(L.ReshapeLayer(agg, (-1, agg.output_shape[-1])), 3, name='out')
        layers = L.get_all_layers(out)
        branch_of, n, trunk_fed = schedule.branch_assignment(layers, {mask})
        assert n == n_streams
        # dependency sets, recomputed independently by a graph walk from every layer
        def deps(l, seen=None):
            seen = set() if seen is None else seen
            if isinstance(l, L.InputLayer):
                return set() if l is mask else {l}
            r = set()
            for i in _ins(l):
                r |= deps(i)
            return r
        for l in layers:
            d = deps(l)
            assert (branch_of[l] is not None) == (len(d) == 1), l.name
        sib, order = schedule.lstm_sibling_groups(layers, branch_of)
        pos = {id(l): k for k, l in enumerate(order)}
        assert len(pos) == len(layers)
        for l in layers:
            for i in _ins(l):
                assert pos[id(l)] < pos[id(i)]
        for l, others in sib.items():
            for o in others:
                assert o.input_layers[0] is l.input_layers[0] and l not in _ins(o) and o not in _ins(l)
        if blstm:
            assert sib.get(f) == (b,) and pos[id(f)] < pos[id(b)]

    run()
