"""Synthetic small instances of every modelzoo builder + feeds, shared by CPU (wiring) and GPU (parity) tests."""
import numpy as np

from ipavsr_b200 import layers as L
from ipavsr_b200 import modelzoo, nonlinearities as nl, init
from ipavsr_b200.function import tensor as T


class FakeDBN(object):
    """The legacy encoder object form: get_all_layers()[1..4] carry .W/.b (modelzoo/deltanet.py:63-73)."""

    class _Lyr(object):
        def __init__(self, W, b):
            self.W, self.b = W, b

    def __init__(self, weights, biases):
        self._layers = [None] + [FakeDBN._Lyr(w, b) for w, b in zip(weights, biases)]

    def get_all_layers(self):
        return self._layers


def enc_weights(rng, D, shapes=(2000, 1000, 500, 50)):
    s = [D] + list(shapes)
    W = [rng.normal(0, 1.0 / np.sqrt(s[i]), (s[i], s[i + 1])).astype('float32') for i in range(len(shapes))]
    b = [rng.normal(0, 0.1, (s[i + 1],)).astype('float32') for i in range(len(shapes))]
    return W, b


def ae_tuple(rng, D, shapes=(60, 40, 30, 10), acts=('sigmoid', 'rectify', 'tanh', 'linear')):
    W, b = enc_weights(rng, D, shapes)
    return W, b, list(shapes), [nl.select_nonlinearity(a) for a in acts]


def make_feed(rng, N, T, dims, lens=None):
    if lens is None:
        lens = rng.integers(2, T + 1, size=N)
        lens[0] = T
    mask = (np.arange(T)[None, :] < np.asarray(lens)[:, None]).astype('uint8')
    xs = []
    for D in dims:
        x = rng.normal(size=(N, T, D)).astype('float32')
        x *= mask[:, :, None]            # zero padding past len (utils/datagen.py:138-139)
        xs.append(x)
    return xs, mask, lens


def build(name, rng, C=7, H=12, win=3, fusiontype='sum'):
    """Returns dict(net, inputs=[(layer-name, array-index)], level='seq'|'frame', dims=[...], extra)."""
    np.random.seed(int(rng.integers(1 << 30)))
    v = lambda n: T.tensor3(n)
    m = T.matrix('mask', dtype='uint8')
    if name == 'deltanet':
        D = 40
        W, b = enc_weights(rng, D)
        net = modelzoo.deltanet.create_model(FakeDBN(W, b), (None, None, D), v('x'), (None, None), m, H, win, C)
        return dict(net=net, names=['input'], dims=[D], level='seq')
    if name == 'deltanet_majority_vote':
        D = 40
        net = modelzoo.deltanet_majority_vote.create_model(ae_tuple(rng, D), (None, None, D), v('x'), (None, None), m,
                                                           H, win, C, init.GlorotUniform(), True, True)
        return dict(net=net, names=['input'], dims=[D], level='frame')
    if name == 'deltanet_v1':
        D = 18
        net = modelzoo.deltanet_v1.create_model((None, None, D), v('x'), (None, None), m, win, H, C,
                                                init.GlorotUniform(), False, False)
        return dict(net=net, names=['input'], dims=[D], level='frame')
    if name == 'lstm_classifier_baseline':
        D = 30
        net = modelzoo.lstm_classifier_baseline.create_model((None, None, D), v('x'), (None, None), m, H, C)
        return dict(net=net, names=['input'], dims=[D], level='seq')
    if name == 'adenet_v1':
        D, Dd = 40, 18
        W, b = enc_weights(rng, D)
        net, _ = modelzoo.adenet_v1.create_model(FakeDBN(W, b), (None, None, D), v('x'), (None, None), m,
                                                 (None, None, Dd), v('dct'), H, win, C)
        return dict(net=net, names=['input', 'dct'], dims=[D, Dd], level='seq')
    if name == 'adenet_v2':
        D, Dd = 40, 18
        net, fuse = modelzoo.adenet_v2.create_model(ae_tuple(rng, D), (None, None, D), v('x'), (None, None), m,
                                                    (None, None, Dd), v('dct'), H, win, C, fusiontype,
                                                    init.GlorotUniform(), True)
        return dict(net=net, names=['input', 'dct'], dims=[D, Dd], level='frame', fuse=fuse)
    if name == 'adenet_v3':
        D, Dd = 40, 18
        W1, b1 = enc_weights(rng, D)
        W2, b2 = enc_weights(rng, D)
        net, fuse = modelzoo.adenet_v3.create_model(FakeDBN(W1, b1), FakeDBN(W2, b2), (None, None, D), v('x'),
                                                    (None, None), m, (None, None, Dd), v('dct'), (None, None, D),
                                                    v('diff'), H, win, C, fusiontype)
        return dict(net=net, names=['raw_im', 'dct', 'diff_im'], dims=[D, Dd, D], level='seq', fuse=fuse)
    if name == 'adenet_3stream':
        dims = [40, 24, 32]
        aes = [ae_tuple(rng, d) for d in dims]
        net, fuse = modelzoo.adenet_3stream.create_model(aes[0], aes[1], aes[2], (None, None, dims[0]), v('s1'),
                                                         (None, None, dims[1]), v('s2'), (None, None, dims[2]),
                                                         v('s3'), (None, None), m, H, win, C, fusiontype)
        return dict(net=net, names=['s1_im', 's2_im', 's3_im'], dims=dims, level='frame', fuse=fuse)
    if name == 'adenet_4stream':
        dims = [40, 24, 32, 20]
        aes = [ae_tuple(rng, d) for d in dims]
        net, fuse = modelzoo.adenet_4stream.create_model(aes[0], aes[1], aes[2], aes[3],
                                                         (None, None, dims[0]), v('s1'), (None, None, dims[1]), v('s2'),
                                                         (None, None, dims[2]), v('s3'), (None, None, dims[3]), v('s4'),
                                                         (None, None), m, H, win, C, fusiontype)
        return dict(net=net, names=['s1_im', 's2_im', 's3_im', 's4_im'], dims=dims, level='frame', fuse=fuse)
    if name in VARIANTS:
        return _build_variant(name, rng, C, H, win, fusiontype, v, m)
    raise KeyError(name)


def lstm_mat(rng, I, H, prefixes=('f_lstm', 'b_lstm')):
    """An LSTM `.mat` dict in the layout of `modelzoo/deltanet_majority_vote.py:183-194`."""
    d = {}
    for pre in prefixes:
        for g in ('ingate', 'forgetgate', 'cell', 'outgate'):
            d['%s_w_in_to_%s' % (pre, g)] = rng.normal(0, 0.2, (I, H)).astype('float32')
            d['%s_w_hid_to_%s' % (pre, g)] = rng.normal(0, 0.2, (H, H)).astype('float32')
            d['%s_b_%s' % (pre, g)] = rng.normal(0, 0.1, (1, H)).astype('float32')
    return d


def _build_variant(name, rng, C, H, win, fusiontype, v, m):
    """the remaining builder variants (same layers, different wiring)."""
    Z = modelzoo
    D, Dd = 40, 18
    sh = lambda d: (None, None, d)
    ms = (None, None)
    dbn = lambda: FakeDBN(*enc_weights(rng, D))
    rect = ('rectify', 'rectify', 'rectify', 'linear')
    if name == 'adenet_v1_1':
        net = Z.adenet_v1_1.create_model(dbn(), sh(D), v('x'), ms, m, sh(Dd), v('dct'), H, win, C)
        return dict(net=net, names=['input', 'dct'], dims=[D, Dd], level='seq')
    if name == 'adenet_v2_1':
        net, fuse = Z.adenet_v2_1.create_model(dbn(), dbn(), sh(D), v('x'), ms, m, sh(D), v('diff'), H, win, C, fusiontype)
        return dict(net=net, names=['raw_im', 'diff_im'], dims=[D, D], level='seq', fuse=fuse)
    if name in ('adenet_v2_2', 'adenet_v2_nodelta', 'adenet_v2_4', 'adenet_2stream', 'adenet_2stream_pretrained',
                'adenet_2stream_pretrained_blstm'):
        dims = [D, 24]
        aes = [ae_tuple(rng, d, acts=rect) for d in dims]
        if name == 'adenet_v2_2':
            net, fuse = Z.adenet_v2_2.create_model(aes[0], aes[1], sh(dims[0]), v('s1'), ms, m, sh(dims[1]), v('s2'), H,
                                                   win, C, fusiontype)
        elif name == 'adenet_v2_nodelta':
            net, fuse = Z.adenet_v2_nodelta.create_model(aes[0], aes[1], sh(dims[0]), v('s1'), ms, m, sh(dims[1]), v('s2'),
                                                         H, C, fusiontype)
        elif name == 'adenet_v2_4':
            net, fuse = Z.adenet_v2_4.create_model(aes[0], aes[1], sh(dims[0]), v('s1'), ms, m, sh(dims[1]), v('s2'), H,
                                                   win, C, fusiontype)
            return dict(net=net, names=['raw_im', 'diff_im'], dims=dims, level='frame', fuse=fuse)
        elif name == 'adenet_2stream':
            net, fuse = Z.adenet_2stream.create_model(aes[0], aes[1], sh(dims[0]), v('s1'), sh(dims[1]), v('s2'), ms, m, H,
                                                      win, C, fusiontype)
        else:
            mats = [lstm_mat(rng, 30, H), lstm_mat(rng, 30, H)]
            net, fuse = Z.adenet_2stream.create_pretrained_model(
                aes[0], mats[0], aes[1], mats[1], sh(dims[0]), v('s1'), sh(dims[1]), v('s2'), ms, m, H, win, C, fusiontype,
                init.Orthogonal(), True, name.endswith('blstm'))
            return dict(net=net, names=['s1_im', 's2_im'], dims=dims, level='frame', fuse=fuse, mats=mats)
        return dict(net=net, names=['s1_im', 's2_im'], dims=dims, level='frame', fuse=fuse)
    if name == 'adenet_v2_3':
        net, fuse = Z.adenet_v2_3.create_model(dbn(), sh(D), v('x'), ms, m, sh(Dd), v('dct'), H, win, C, fusiontype)
        return dict(net=net, names=['input', 'dct'], dims=[D, Dd], level='frame', fuse=fuse)
    if name == 'adenet_v4':
        net, fuse = Z.adenet_v4.create_model(dbn(), sh(D), v('x'), ms, m, sh(Dd), v('dct'), H, win, C)
        return dict(net=net, names=['input', 'dct'], dims=[D, Dd], level='seq', fuse=fuse)
    if name == 'adenet_v5':
        net, fuse = Z.adenet_v5.create_model(dbn(), dbn(), sh(D), v('x'), ms, m, sh(Dd), v('dct'), sh(D), v('diff'), H, win,
                                             C, fusiontype == 'adasum')
        return dict(net=net, names=['raw_im', 'dct', 'diff_im'], dims=[D, Dd, D], level='seq', fuse=fuse)
    if name == 'adenet_v6':
        net, fuse = Z.adenet_v6.create_model(dbn(), dbn(), sh(D), v('x'), ms, m, sh(D), v('diff'), H, win, C,
                                             fusiontype == 'adasum')
        return dict(net=net, names=['raw_im', 'diff_im'], dims=[D, D], level='seq', fuse=fuse)
    if name in ('adenet_3stream_dct', 'adenet_3stream_dropout', 'adenet_3stream_pretrained'):
        dims = [D, 24, 32]
        aes = [ae_tuple(rng, d, acts=rect) for d in dims]
        shapes_vars = [sh(dims[0]), v('s1'), sh(dims[1]), v('s2'), sh(dims[2]), v('s3')]
        if name == 'adenet_3stream_dct':
            net, fuse = Z.adenet_3stream_dct.create_model(aes[0], aes[1], *(shapes_vars + [ms, m, H, win, C, fusiontype]))
        elif name == 'adenet_3stream_dropout':
            net, fuse = Z.adenet_3stream_dropout.create_model(aes[0], aes[1], aes[2],
                                                              *(shapes_vars + [ms, m, H, win, C, fusiontype]))
        else:
            mats = [lstm_mat(rng, 30, H, ('f_lstm',)) for _ in dims]
            net, fuse = Z.adenet_3stream.create_pretrained_model(aes[0], mats[0], aes[1], mats[1], aes[2], mats[2],
                                                                 *(shapes_vars + [ms, m, H, win, C, fusiontype]))
            return dict(net=net, names=['s1_im', 's2_im', 's3_im'], dims=dims, level='frame', fuse=fuse, mats=mats)
        return dict(net=net, names=['s1_im', 's2_im', 's3_im'], dims=dims, level='frame', fuse=fuse)
    if name == 'avnet':
        dims = [D, 24]
        subs = []
        for k, d in enumerate(dims):
            W, b = enc_weights(rng, d)
            subs.append(Z.avnet.create_pretrained_substream(W, b, sh(d), v('s%d' % k), ms, m, ('visual', 'audio')[k], H, win))
        net, fuse = Z.avnet.create_model(subs, ms, m, H, C, fusiontype)
        return dict(net=net, names=['input_visual', 'input_audio'], dims=dims, level='frame', fuse=fuse)
    if name in ('lstm_classifier_majority_vote', 'lstm_classifier_majority_vote_lstm'):
        net = Z.lstm_classifier_majority_vote.create_model(sh(Dd), v('x'), ms, m, H, C, init.GlorotUniform(), True,
                                                           not name.endswith('_lstm'))
        return dict(net=net, names=['input'], dims=[Dd], level='frame')
    raise KeyError(name)


VARIANTS = ['adenet_v1_1', 'adenet_v2_1', 'adenet_v2_2', 'adenet_v2_3', 'adenet_v2_4', 'adenet_v2_nodelta', 'adenet_v4',
            'adenet_v5', 'adenet_v6', 'adenet_2stream', 'adenet_2stream_pretrained', 'adenet_2stream_pretrained_blstm',
            'adenet_3stream_dct', 'adenet_3stream_dropout', 'adenet_3stream_pretrained', 'avnet', 'lstm_classifier_majority_vote',
            'lstm_classifier_majority_vote_lstm']

ALL = ['deltanet', 'deltanet_majority_vote', 'deltanet_v1', 'lstm_classifier_baseline', 'adenet_v1', 'adenet_v2',
       'adenet_v3', 'adenet_3stream', 'adenet_4stream']


def randomize_params(net, rng):
    """Give every parameter a non-trivial value (biases, peepholes, inits, BN stats, adacoeffs are otherwise 0/1)."""
    for p in L.get_all_params(net):
        if p.name and (p.name.endswith('.W') and p.shape[0] > 100):
            continue                               # keep the scaled encoder weights
        if p.name and p.name.endswith('inv_std'):
            p.set_value(rng.uniform(0.5, 1.5, p.shape).astype('float32'))
        elif p.name and p.name.startswith('adacoeff'):
            p.set_value(np.float32(rng.uniform(0.5, 1.5)))
        elif len(p.shape) == 2 and p.shape[0] > 1:
            p.set_value((p.get_value() + rng.normal(0, 0.05, p.shape)).astype('float32'))
        else:
            p.set_value(rng.normal(0, 0.3, p.shape).astype('float32'))


def input_layers(net):
    return {l.name: l for l in L.get_all_layers(net) if isinstance(l, L.InputLayer)}


def dropout_masks_for(net, rng, N, T):
    out = {}
    for l in L.get_all_layers(net):
        if isinstance(l, L.DropoutLayer):
            F = l.output_shape[-1]
            out[l.name] = (rng.random((N, T, F)) >= l.p).astype('uint8')
    return out


def rectify_aligner(net, run, N, T):
    """Returns (callback for OracleNet.loss_and_grads(after_forward=...), flips list).

    A rectify unit whose pre-activation lies within float32 rounding of 0 can take the other branch on the device than in
    the float64 oracle; its sub-gradient then differs by the whole upstream gradient of that (row, unit), which moves one
    column of the layer's weight gradient by the contribution of one frame.  The callback counts such units, REQUIRES each
    of them to be at rounding distance from 0 (|z| < 1e-5 max|z|: anything else is a real error and fails the test), and
    makes the oracle's backward take the device's branch there — every gradient must then agree to the normal gate."""
    rect = [l for l in L.get_all_layers(net) if isinstance(l, L.DenseLayer) and l.nonlinearity.name == 'rectify']
    flips = []

    def align(orc):
        for l in rect:
            saved = run.saved.get(l)
            yd = saved if (saved is not None and not isinstance(saved, tuple)) else run.vals[l][0]
            ydev = yd.torch_view().cpu().numpy() > 0
            plan = getattr(run, 'plan', None)
            if plan is not None and ydev.shape[0] == plan.M + 1 and plan.M + 1 != N * T:
                pk = dict(plan.tables)['pack']                     # packed rows -> padded rows in the caller's order
                full = np.empty((N * T, ydev.shape[1]), bool)
                full[:] = ydev[plan.M]                             # padding frames share the zero row's output
                full[pk[:-1]] = ydev[:-1]
                ydev = full
            elif plan is not None:
                ydev = ydev[dict(plan.tables)['unperm']]
            x, z, yo = orc.caches[l]
            diff = ydev != (z > 0)
            if diff.any():
                zmax = np.abs(z).max()
                assert np.abs(z[diff]).max() < 1e-5 * zmax, ('a flipped rectify unit is not at rounding distance from 0',
                                                           l.name, np.abs(z[diff]).max(), zmax)
                flips.append((l.name, int(diff.sum())))
                z = z.copy()
                z[diff] = np.where(ydev[diff], 1e-300, -1e-300)
                orc.caches[l] = (x, z, yo)
        total = sum(N * T * l.num_units for l in rect)
        assert sum(n for _, n in flips) <= max(8, 1e-5 * total), flips

    return align, flips


def grad_errors(params, grads, grads_ref):
    """[(param, error / scale)]: max abs error of every gradient against max(its own max, 2e-2 of the largest gradient of
    the network); a gradient that is analytically zero (the bias in front of a BatchNormLayer: the float64 oracle returns
    ~1e-18) is pure cancellation noise of a sum over all rows and is compared against the largest gradient instead."""
    gmax = max(np.abs(gr).max() for gr in grads_ref)
    out = []
    for p, g, gr in zip(params, grads, grads_ref):
        own = np.abs(gr).max()
        scale = gmax if own < 1e-10 * gmax else max(own, 2e-2 * gmax)
        out.append((p, np.abs(g - gr).max() / scale))
    return out
