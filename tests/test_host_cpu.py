"""CPU-side tests: the C-ABI library loads and exports every declared symbol (no compute), builder wiring and
parameter order, error conventions, oracle network gradients vs finite differences."""
import ctypes
import os
import re

import numpy as np
import pytest

from ipavsr_b200 import layers as L, _lib, nonlinearities as nl
from ipavsr_b200.custom.updates import generate_lr_map
import model_util as MU


def test_library_builds_loads_and_exports_every_header_symbol():
    lib = _lib.load()
    names = _lib.header_symbols()
    assert len(names) >= 30
    for n in names:
        assert hasattr(lib, n), n
        assert n in _lib._SIGS, 'no ctypes signature for %s' % n
    assert lib.ipavsr_version() >= 100
    assert isinstance(lib.ipavsr_launch_count(), int)


def test_ctypes_arity_matches_header():
    text = re.sub(r'/\*.*?\*/', '', open(_lib.HEADER).read(), flags=re.S)
    for m in re.finditer(r'\b(ipavsr_[a-z0-9_]+)\s*\(([^;]*?)\)\s*;', text, flags=re.S):
        name, args = m.group(1), m.group(2).strip()
        n = 0 if args in ('', 'void') else args.count(',') + 1
        assert len(_lib._SIGS[name][1]) == n, name


def test_argument_errors_do_not_need_a_gpu():
    lib = _lib.load()
    rc = lib.ipavsr_gemm(0, 0, 0, -1, 4, 4, None, 4, None, 4, None, 4, None, 0, 0, None, 0, None)
    assert rc == -1 and b'ipavsr_gemm' in lib.ipavsr_last_error()
    rc = lib.ipavsr_delta_fwd(None, 4, None, 12, 1, 1, 4, 1, 1, None)
    assert rc == -1
    # the entry points of SURVEY 8f ranks 3 and 4: bad arguments are refused before anything touches the device
    one = ctypes.c_void_p(16)                       # a non-NULL, never dereferenced pointer
    assert lib.ipavsr_dct_basis(None, 4, None, 8, 4, None) == -1
    assert lib.ipavsr_dct_basis(one, 2, None, 8, 4, None) == -1 and b'ldb' in lib.ipavsr_last_error()
    assert lib.ipavsr_dct_project(one, 8, one, 4, one, 2, 10, 8, 4, None) == -1          # ldo < K
    assert lib.ipavsr_dct_project(one, 8, one, 4, one, 4, 65536 * 128, 8, 4, None) == -1  # more than 65535 frame tiles
    assert lib.ipavsr_reorder(one, 12, one, 12, 5, 3, 4, 1, None) == -1                  # in place is not supported
    assert lib.ipavsr_reorder(one, 8, ctypes.c_void_p(32), 12, 5, 3, 4, 1, None) == -1   # ldx < d1*d2
    assert lib.ipavsr_align_fill(one, 8, ctypes.c_void_p(32), 8, None, one, None, 3, 8, 10, None) == -1
    assert lib.ipavsr_gather_cols(one, 8, one, one, 2, 10, 4, None) == -1                # ldo < K
    assert lib.ipavsr_col_abs_sum(one, 2, one, 10, 4, None) == -1                        # ldx < F
    assert lib.ipavsr_squared_error(one, 4, one, 8, one, None, 0, 10, 8, 1.0, None) == -1   # ldp < F
    assert lib.ipavsr_l2_penalty(one, None, 100, None, one, one, 1.0, None) == -1
    assert lib.ipavsr_zigzag_indices(0, 4, None) == -1
    # empty problems are a no-op that needs no device either
    assert lib.ipavsr_dct_project(one, 8, one, 4, one, 4, 0, 8, 4, None) == 0
    assert lib.ipavsr_align_fill(one, 8, ctypes.c_void_p(32), 8, one, one, None, 3, 8, 0, None) == 0


def test_adenet_v2_layer_and_param_order():
    rng = np.random.default_rng(0)
    spec = MU.build('adenet_v2', rng, fusiontype='adasum')
    names = [l.name for l in L.get_all_layers(spec['net'])]
    assert names == ['input', 'reshape1', 'fc1', 'fc2', 'fc3', 'bottleneck', 'reshape2', 'delta', 'mask', 'lstm_bn',
                     'dct', 'delta_dct', 'lstm_dct', 'adasum', 'f_lstm_agg', 'b_lstm_agg', 'sum2', 'reshape3',
                     'softmax', 'output']
    pn = [p.name for p in L.get_all_params(spec['net'])]
    assert pn[:8] == ['fc1.W', 'fc1.b', 'fc2.W', 'fc2.b', 'fc3.W', 'fc3.b', 'bottleneck.W', 'bottleneck.b']
    lstm = ['W_in_to_ingate', 'W_hid_to_ingate', 'b_ingate', 'W_in_to_forgetgate', 'W_hid_to_forgetgate',
            'b_forgetgate', 'W_in_to_cell', 'W_hid_to_cell', 'b_cell', 'W_in_to_outgate', 'W_hid_to_outgate',
            'b_outgate']
    peep = ['W_cell_to_ingate', 'W_cell_to_forgetgate', 'W_cell_to_outgate']
    init = ['cell_init', 'hid_init']
    expect = ['lstm_bn.' + n for n in lstm + peep + init] + ['lstm_dct.' + n for n in lstm + peep + init] + \
             ['adacoeff0', 'adacoeff1'] + ['f_lstm_agg.' + n for n in lstm + init] + \
             ['b_lstm_agg.' + n for n in lstm + init] + ['softmax.W', 'softmax.b']
    assert pn[8:] == expect            # the agg BLSTM never gets peepholes (adenet_v2.py:77)
    assert [p.name for p in L.get_all_params(spec['fuse'], scaling_param=True)] == ['adacoeff0', 'adacoeff1']


def test_adenet_v3_order_matches_notebook_print():
    """avletters/avletters_training.ipynb:462-492 prints this traversal for adenet_v3."""
    rng = np.random.default_rng(1)
    spec = MU.build('adenet_v3', rng, H=10)
    names = [l.name for l in L.get_all_layers(spec['net'])]
    assert names == ['raw_im', 'reshape1_raw', 'fc1_raw', 'fc2_raw', 'fc3_raw', 'bottleneck_raw', 'reshape2_raw',
                     'delta_raw', 'dropout_raw', 'mask', 'lstm_raw', 'dct', 'dropout_dct', 'lstm_dct', 'diff_im',
                     'reshape1_diff', 'fc1_diff', 'fc2_diff', 'fc3_diff', 'bottleneck_diff', 'reshape2_diff',
                     'delta_diff', 'dropout_diff', 'lstm_diff', 'sum1', 'dropout_agg', 'f_lstm_agg', 'b_lstm_agg',
                     'sum2', 'slice1', 'output']
    lstm_raw = [l for l in L.get_all_layers(spec['net']) if l.name == 'lstm_raw'][0]
    assert lstm_raw.num_units == 20 and lstm_raw.peepholes          # int(lstm_size/(1-0.5)), Lasagne default peepholes
    agg = [l for l in L.get_all_layers(spec['net']) if l.name == 'f_lstm_agg'][0]
    assert agg.num_units == 20 and agg.peepholes


@pytest.mark.parametrize('name', MU.ALL)
def test_every_builder_builds(name):
    rng = np.random.default_rng(2)
    spec = MU.build(name, rng, fusiontype='concat' if name.startswith('adenet_') and name != 'adenet_v1' else 'sum')
    net = spec['net']
    shapes = net.output_shape
    assert shapes[-1] == 7
    assert (len(shapes) == 3) == (spec['level'] == 'frame')
    vals = L.get_all_param_values(net)
    L.set_all_param_values(net, vals)
    with pytest.raises(ValueError):
        L.set_all_param_values(net, vals[:-1])


def test_shapes_and_quirks():
    rng = np.random.default_rng(3)
    s = MU.build('adenet_v1', rng, H=8)
    byname = {l.name: l for l in L.get_all_layers(s['net'])}
    assert byname['concat'].output_shape[-1] == 150 + 18
    assert byname['f_lstm1'].peepholes and byname['f_lstm2'].num_units == 16
    assert [p.name for p in byname['batchnorm1'].params] == ['batchnorm1.beta', 'batchnorm1.gamma', 'batchnorm1.mean',
                                                             'batchnorm1.inv_std']
    assert [p.name for p in byname['batchnorm1'].get_params(trainable=True)] == ['batchnorm1.beta', 'batchnorm1.gamma']
    s2 = MU.build('adenet_v2', rng, fusiontype='concat')
    b2 = {l.name: l for l in L.get_all_layers(s2['net'])}
    assert b2['delta_dct'].output_shape[-1] == 54 and b2['f_lstm_agg'].num_inputs == 24 and not b2['f_lstm_agg'].peepholes
    s4 = MU.build('adenet_4stream', rng, fusiontype='concat')
    assert {l.name: l for l in L.get_all_layers(s4['net'])}['f_lstm_agg'].num_inputs == 48
    with pytest.raises(ValueError):
        MU.build('adenet_v2', rng, fusiontype='bogus')
    with pytest.raises(KeyError):
        nl.select_nonlinearity('nope')


def test_generate_lr_map_prefix_rule():
    rng = np.random.default_rng(4)
    s = MU.build('adenet_v2', rng, fusiontype='adasum')
    params = L.get_all_params(s['net'], trainable=True)
    m = generate_lr_map(params, {'fc1': 0.5, 'lstm_bn': 0.25, 'adacoeff': 0.125}, 1.0)
    byname = {p.name: v for p, v in m.items()}
    assert byname['fc1.W'] == 0.5 and byname['fc2.W'] == 1.0 and byname['lstm_bn.cell_init'] == 0.25
    assert byname['adacoeff0'] == 0.125          # 'adacoeff0'[:rfind('.')] == 'adacoeff' (custom/updates.py:27)


@pytest.mark.parametrize('name', ['adenet_v2', 'adenet_v1', 'adenet_v3'])
def test_oracle_network_gradients_vs_finite_differences(name):
    from oracle.net import OracleNet
    rng = np.random.default_rng(7)
    spec = MU.build(name, rng, C=5, H=6, win=2, fusiontype='adasum')
    net = spec['net']
    MU.randomize_params(net, rng)
    N, T = 4, 6
    xs, mask, lens = MU.make_feed(rng, N, T, spec['dims'])
    feed = dict(zip(spec['names'], xs))
    feed['mask'] = mask
    y1 = rng.integers(0, 5, size=N)
    level = spec['level']
    y = y1 if level == 'seq' else np.repeat(y1[:, None], T, 1)
    loss_name = 'categorical_crossentropy' if level == 'seq' else 'temporal_softmax'
    dm = MU.dropout_masks_for(net, rng, N, T)
    o = OracleNet(net, np.float64)
    f = lambda: float(o.loss_and_grads(feed, 2, y, mask, loss_name, False, dm, update_bn=False)[0])
    _, _, grads = o.loss_and_grads(feed, 2, y, mask, loss_name, False, dm, update_bn=False)
    params = L.get_all_params(net, trainable=True)
    eps = 1e-5
    checked = 0
    for p, g in zip(params, grads):
        base = p.get_value().astype(np.float64)
        flat = max(int(np.prod(p.shape)), 1)
        idx = np.unravel_index(int(rng.integers(0, flat)), p.shape) if p.shape else ()
        vals = {}
        for sgn in (1, -1):
            v = base.copy()
            v[idx] = v[idx] + sgn * eps
            p.get_value = (lambda vv: (lambda: vv))(v)        # float64 value straight into the float64 oracle
            vals[sgn] = f()
            del p.get_value
        num = (vals[1] - vals[-1]) / (2 * eps)
        ana = float(g[idx])
        assert abs(num - ana) < 1e-6 + 1e-4 * abs(num), (p.name, idx, num, ana)
        checked += 1
    assert checked == len(params)


def test_zigzag_index_walk_of_the_c_abi_matches_the_reference_vectors():
    """ipavsr_zigzag_indices is host-side index work: checked here without a GPU against the reference's own outputs."""
    import ctypes as C
    FG = np.load(os.path.join(os.path.dirname(__file__), 'golden', 'features.npz'))
    lib = _lib.load()
    for r, c in FG['zigzag_shapes']:
        want = FG['zigzag_%d_%d' % (r, c)]
        order = np.empty(int(r) * int(c), dtype=np.int32)
        rc = lib.ipavsr_zigzag_indices(int(r), int(c), order.ctypes.data_as(C.c_void_p))
        if want[0] == -1:
            assert rc == -1
        else:
            assert rc == 0
            np.testing.assert_array_equal(order, want)


def test_force_align_gather_plan_matches_the_reference_vectors(monkeypatch):
    """The host half of force_align / multistream_force_align (offsets, fill rows, target gather, in-place length update)
    with the device row gather replaced by its NumPy definition."""
    from ipavsr_b200.utils import preprocessing as P
    FG = np.load(os.path.join(os.path.dirname(__file__), 'golden', 'features.npz'))

    def gather(x, in_off, out_off, fill):
        x = np.asarray(x)
        u = np.repeat(np.arange(len(in_off) - 1), np.diff(out_off))
        j = np.arange(out_off[-1]) - out_off[u]
        return x[np.where(j < np.diff(in_off)[u], in_off[u] + j, fill[u])]

    monkeypatch.setattr(P, '_align_gather', gather)
    (na, nta, nl1), (nb, ntb, nl2) = P.force_align((FG['fa_a'], FG['fa_ta'], FG['fa_l1'].copy()),
                                                   (FG['fa_b'], FG['fa_tb'], FG['fa_l2'].copy()))
    for got, key in ((na, 'fa_out_a'), (nta, 'fa_out_ta'), (nl1, 'fa_out_l1'), (nb, 'fa_out_b'), (ntb, 'fa_out_tb'),
                     (nl2, 'fa_out_l2')):
        np.testing.assert_array_equal(got, FG[key])
    res = P.multistream_force_align([(FG['fa_a'], FG['fa_ta'], FG['fa_l1'].copy()),
                                     (FG['fa_b'], FG['fa_tb'], FG['fa_l2'].copy()),
                                     (FG['ms_c'], FG['ms_tc'], FG['ms_l3'].copy())])
    for j, (x, t, l) in enumerate(res):
        np.testing.assert_array_equal(x, FG['ms_out_x%d' % j])
        np.testing.assert_array_equal(t, FG['ms_out_t%d' % j])
        np.testing.assert_array_equal(l, FG['ms_out_l%d' % j])
    with pytest.raises(IndexError):
        P.force_align((np.zeros((8, 4), 'float32'), np.zeros(8, 'uint8'), np.array([3, 5])),
                      (np.zeros((5, 4), 'float32'), np.zeros(5, 'uint8'), np.array([3, 2])))


def test_nolearn_train_split_is_the_first_unshuffled_fold():
    from sklearn.model_selection import KFold
    from ipavsr_b200.custom.nolearn_net import train_split
    for n in (5, 7, 100, 103, 700, 1001):
        tr_want, va_want = next(iter(KFold(5).split(np.zeros((n, 1)))))
        tr, va = train_split(n, 0.2)
        np.testing.assert_array_equal(tr, tr_want)
        np.testing.assert_array_equal(va, va_want)


def test_squared_error_l2_objective_expression_and_oracle_gradients():
    """`T.mean(squared_error(out, y)) + 0.005 * regularize_network_params(net, l2)` builds one loss expression; the oracle's
    value and gradients of that objective agree with central finite differences (float64)."""
    from ipavsr_b200.function import tensor as T, LossExpr
    from ipavsr_b200.custom.objectives import squared_error
    from ipavsr_b200.regularization import regularize_network_params, l2
    from oracle.net import OracleNet
    rng = np.random.default_rng(0)
    x = T.tensor3('x')
    l_in = L.InputLayer((None, None, 6), x, name='input')
    l = L.ReshapeLayer(l_in, (-1, 6))
    l = L.DenseLayer(l, 4, W=rng.normal(size=(6, 4)).astype('float32'), b=rng.normal(size=4).astype('float32'),
                     nonlinearity=nl.sigmoid, name='l1')
    out = L.DenseLayer(l, 6, W=rng.normal(size=(4, 6)).astype('float32'), b=rng.normal(size=6).astype('float32'),
                       nonlinearity=nl.linear, name='output')
    y = T.matrix('y')
    loss = T.mean(squared_error(L.get_output(out), y)) + 0.005 * regularize_network_params(out, l2)
    assert isinstance(loss, LossExpr) and loss.kind == 'squared_error' and abs(loss.l2 - 0.005) < 1e-12
    X = rng.normal(size=(5, 1, 6))
    params = L.get_all_params(out, trainable=True)

    def value():
        return float(OracleNet(out, np.float64).loss_and_grads({'input': X}, 0, X.reshape(5, 6), None, 'squared_error',
                                                               l2=0.005)[0])

    _, _, grads = OracleNet(out, np.float64).loss_and_grads({'input': X}, 0, X.reshape(5, 6), None, 'squared_error', l2=0.005)
    for p, g in zip(params, grads):
        w = p.get_value().astype(np.float64)
        idx = tuple(rng.integers(0, s) for s in w.shape)
        eps = 1e-3
        keep = p.get_value().copy()
        wp = keep.copy(); wp[idx] += eps; p.set_value(wp); up = value()
        wm = keep.copy(); wm[idx] -= eps; p.set_value(wm); dn = value()
        p.set_value(keep)
        assert abs((up - dn) / (2 * eps) - g[idx]) <= 2e-3 * max(abs(g[idx]), 1e-3), p.name


@pytest.mark.parametrize('name', MU.VARIANTS)
def test_every_builder_variant_builds_and_runs_through_the_oracle(name):
    """the remaining modelzoo variants are wiring over the same layers; each builds with the reference's
    positional signature, round-trips its parameter list, and its oracle forward gives normalised class probabilities."""
    from oracle.net import OracleNet
    rng = np.random.default_rng(4)
    spec = MU.build(name, rng, fusiontype='sum' if name in ('adenet_v4', 'adenet_v5', 'adenet_v2_3') else 'concat')
    net = spec['net']
    assert net.output_shape[-1] == 7 and (len(net.output_shape) == 3) == (spec['level'] == 'frame')
    vals = L.get_all_param_values(net)
    L.set_all_param_values(net, vals)
    N, T_ = 3, 6
    xs, mask, lens = MU.make_feed(rng, N, T_, spec['dims'])
    feed = dict(zip(spec['names'], xs))
    feed['mask'] = mask
    out = OracleNet(net, np.float64).forward(feed, 3, deterministic=True)
    assert out.shape == ((N, T_, 7) if spec['level'] == 'frame' else (N, 7))
    np.testing.assert_allclose(out.sum(-1), 1.0, rtol=1e-9)


def test_builder_variant_wiring_details():
    rng = np.random.default_rng(5)
    by = lambda spec: {l.name: l for l in L.get_all_layers(spec['net'])}
    # files with a local create_blstm / create_lstm (default use_peepholes=True): the aggregate has peepholes
    for name in ('adenet_v2_2', 'adenet_v2_nodelta', 'adenet_v2_1'):
        assert by(MU.build(name, rng, fusiontype='concat'))['f_lstm_agg'].peepholes, name
    for name in ('adenet_2stream', 'adenet_3stream_dct', 'adenet_3stream_dropout'):     # custom.layers.create_blstm: none
        assert not by(MU.build(name, rng, fusiontype='concat'))['f_lstm_agg'].peepholes, name
    b = by(MU.build('adenet_v2_nodelta', rng, fusiontype='concat'))
    assert not [n for n in b if n and n.startswith('delta')] and b['lstm_s1'].num_inputs == 10
    b = by(MU.build('adenet_v2_4', rng, fusiontype='concat'))
    assert 'b_lstm_agg' not in b and b['f_lstm_agg'].peepholes and b['f_lstm_agg'].num_inputs == 24
    names = [l.name for l in L.get_all_layers(MU.build('adenet_v2_4', rng, fusiontype='sum')['net'])]
    assert names[-4:] == ['f_lstm_agg', None, 'softmax', 'output']
    b = by(MU.build('adenet_3stream_dropout', rng, fusiontype='concat'))
    assert b['lstm_s2'].num_units == 24 and b['f_lstm_agg'].num_units == 24 and b['f_lstm_agg'].num_inputs == 72
    assert b['concat_dropout'].p == 0.5 and b['reshape3'].shape == (-1, 24)
    b = by(MU.build('adenet_3stream_dct', rng, fusiontype='concat'))
    assert 'fc1_s3' not in b and b['delta_s3'].output_shape[-1] == 96 and b['lstm_s3'].num_inputs == 96
    b = by(MU.build('adenet_v4', rng))
    assert b['lstm_bn'].num_units == 24 and b['lstm_bn'].peepholes and b['dropout_dct'].p == 0.2 and 'b_lstm_agg' not in b
    b = by(MU.build('adenet_v6', rng))
    assert 'dct' not in b and 'sum1' in b and b['f_lstm_agg'].num_units == 24
    assert 'adasum1' in by(MU.build('adenet_v5', rng, fusiontype='adasum'))
    b = by(MU.build('adenet_v1_1', rng))
    assert b['f_lstm1'].num_units == 24 and b['dropout1'].input_layer is b['concat'] and b['dropout2'].input_layer is b['sum1']
    # pretrained sub-stream LSTMs: the .mat weights land in the layers; forward + backward are summed per stream
    spec = MU.build('adenet_2stream_pretrained_blstm', rng, fusiontype='concat')
    b = by(spec)
    np.testing.assert_array_equal(b['b_lstm_s2'].W_hid_to_cell.get_value(), spec['mats'][1]['b_lstm_w_hid_to_cell'])
    np.testing.assert_array_equal(b['f_lstm_s1'].W_in_to_ingate.get_value(), spec['mats'][0]['f_lstm_w_in_to_ingate'])
    assert b['b_lstm_s1'].backwards and b['sum_b_lstm_s1'].input_layers == [b['f_lstm_s1'], b['b_lstm_s1']]
    b = by(MU.build('lstm_classifier_majority_vote_lstm', rng))
    assert 'b_lstm' not in b and b['lstm'].peepholes
