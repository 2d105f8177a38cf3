"""Parity of the SHIPPED arithmetic (engine default: f16x3 tensor-core GEMMs + tensor-core LSTM recurrence, packed /
length-sorted execution) on every BASELINE configuration at the reference's own layer sizes, against the float64 oracle:
probabilities 1e-4 with identical argmax, loss, every parameter gradient.

  adenet_3stream concat  — the bench workload (modelzoo/adenet_3stream.py:145), N = 64 train step, N = 512 forward
  adenet_4stream concat  — BASELINE config 5 (modelzoo/adenet_4stream.py:12), peepholes
  adenet_v3 sum          — README-era trimodal (modelzoo/adenet_v3.py:64): 2 x lstm_size = 500-wide LSTMs, supplied dropout masks
  adenet_v1              — OuluVS bimodal (modelzoo/adenet_v1.py:47): D = 1144 + DCT 90, BatchNorm, BLSTM-250 -> BLSTM-500, C = 10
  deltanet               — AVLetters unimodal (modelzoo/deltanet.py:59): rectify encoder, the reference batch N = 26
Run with -m gpu on the B200."""
import numpy as np
import pytest

from ipavsr_b200 import layers as L, modelzoo, init
from ipavsr_b200.engine import Engine
from ipavsr_b200.function import tensor as T
from oracle.net import OracleNet
from oracle import ops
import model_util as MU

pytestmark = pytest.mark.gpu

ENC = (2000, 1000, 500, 50)
SIG = ('sigmoid', 'sigmoid', 'sigmoid', 'linear')
# gradient gate: 1e-4 of the per-tensor max (measured: a few 1e-6); analytically-zero gradients are compared against the
# scale of the others
GRAD_TOL = 1e-4


def _build(name, rng):
    sh = lambda d: (None, None, d)
    m, w = T.matrix('mask', dtype='uint8'), 9
    v = lambda n: T.tensor3(n)
    ae = lambda D, acts=SIG: MU.ae_tuple(rng, D, shapes=ENC, acts=acts)
    dbn = lambda D: MU.FakeDBN(*MU.enc_weights(rng, D, ENC))
    if name == 'adenet_3stream':
        dims = [1200, 1200, 90]
        net, _ = modelzoo.adenet_3stream.create_model(ae(1200), ae(1200), ae(90), sh(1200), v('s1'), sh(1200), v('s2'), sh(90),
                                                      v('s3'), (None, None), m, 250, w, 26, 'concat', init.Orthogonal(), True)
        return net, ['s1_im', 's2_im', 's3_im'], dims, 'frame', 26
    if name == 'adenet_4stream':
        dims = [1200] * 4
        net, _ = modelzoo.adenet_4stream.create_model(ae(1200), ae(1200), ae(1200), ae(1200), sh(1200), v('s1'), sh(1200),
                                                      v('s2'), sh(1200), v('s3'), sh(1200), v('s4'), (None, None), m, 250, w,
                                                      26, 'concat', init.Orthogonal(), True)
        return net, ['s1_im', 's2_im', 's3_im', 's4_im'], dims, 'frame', 26
    if name == 'adenet_v3':
        dims = [1200, 90, 1200]
        net, _ = modelzoo.adenet_v3.create_model(dbn(1200), dbn(1200), sh(1200), v('x'), (None, None), m, sh(90), v('dct'),
                                                 sh(1200), v('diff'), 250, w, 26, 'sum')
        return net, ['raw_im', 'dct', 'diff_im'], dims, 'seq', 26
    if name == 'adenet_v1':
        dims = [1144, 90]
        net, _ = modelzoo.adenet_v1.create_model(dbn(1144), sh(1144), v('x'), (None, None), m, sh(90), v('dct'), 250, w, 10)
        return net, ['input', 'dct'], dims, 'seq', 10
    if name == 'deltanet':
        dims = [1200]
        net = modelzoo.deltanet.create_model(dbn(1200), sh(1200), v('x'), (None, None), m, 250, w, 26)
        return net, ['input'], dims, 'seq', 26
    raise KeyError(name)


def _feed(rng, names, dims, N, Tn, level, C, lo=12):
    lens = rng.integers(lo, Tn + 1, size=N)
    lens[N // 3] = Tn
    xs, mask, _ = MU.make_feed(rng, N, Tn, dims, lens=lens)
    y1 = rng.integers(0, C, size=N).astype('int32')
    y = y1 if level == 'seq' else np.repeat(y1[:, None], Tn, 1).astype('int32')
    feed = dict(zip(names, xs))
    feed['mask'] = mask
    return feed, mask, y


def _rectify_layers(net):
    return [l for l in L.get_all_layers(net) if isinstance(l, L.DenseLayer) and l.nonlinearity.name == 'rectify']


def _check(name, N, seed, train=True, packed=None):
    rng = np.random.default_rng(seed)
    np.random.seed(seed)
    net, names, dims, level, C = _build(name, rng)
    MU.randomize_params(net, rng)
    feed, mask, y = _feed(rng, names, dims, N, 40, level, C)
    dm = MU.dropout_masks_for(net, rng, N, 40)
    lname = 'categorical_crossentropy' if level == 'seq' else 'temporal_softmax'
    eng = Engine(net, packed=packed)                     # engine defaults: f16x3 + tensor-core LSTM
    assert eng.gemm_mode == 4
    ins = MU.input_layers(net)
    run, out = eng.forward({ins[k]: v for k, v in feed.items()}, 9, deterministic=not train, train=train, dropout_masks=dm,
                           update_bn=False)
    probs = eng.read(out)
    o = OracleNet(net, np.float64)
    if not train:
        ref = o.forward(feed, 9, deterministic=True)
        rel = np.abs(probs.reshape(ref.shape) - ref).max() / np.abs(ref).max()
        assert rel < 1e-4, (name, N, rel)
        assert (probs.reshape(ref.shape).argmax(-1) == ref.argmax(-1)).all()
        return rel, None, 0
    # rectify units whose float64 pre-activation is within float32 rounding of 0 may take the other branch on the device:
    # counted, required to be at rounding distance, and the oracle's backward takes the device's branch there
    # (model_util.rectify_aligner); everything must then agree to GRAD_TOL
    align, flips = MU.rectify_aligner(net, run, N, 40)
    rect = _rectify_layers(net)
    loss_ref, out_ref, grads_ref = o.loss_and_grads(feed, 9, y, mask, lname, deterministic=False, dropout_masks=dm,
                                                    update_bn=False, after_forward=align)
    nflip = sum(n for _, n in flips)
    total = sum(N * 40 * l.num_units for l in rect)
    rel = np.abs(probs.reshape(out_ref.shape) - out_ref).max() / np.abs(out_ref).max()
    assert rel < 1e-4, (name, N, 'probs', rel)
    assert (probs.reshape(out_ref.shape).argmax(-1) == out_ref.argmax(-1)).all()
    eng.loss_and_backward(run, out, lname, y, mask, count=float(mask.sum()))
    loss = eng.read_loss()
    assert abs(loss - loss_ref) < 1e-4 * abs(loss_ref), (name, loss, loss_ref)
    params = L.get_all_params(net, trainable=True)
    grads = eng.param_grads(params)
    errs = MU.grad_errors(params, grads, grads_ref)
    for p, err in errs:
        assert err < GRAD_TOL, (name, N, p.name, err, flips)
    worst = max(e for _, e in errs)
    print('PARITY %s N=%d reference sizes (engine defaults): probs rel %.2e, worst grad err %.2e, rectify flips %d of %d' % (name, N, rel, worst, nflip, total))
    return rel, worst, nflip


@pytest.mark.parametrize('name,N', [('adenet_3stream', 64), ('adenet_4stream', 32), ('adenet_v3', 26), ('adenet_v1', 10),
                                    ('adenet_v1', 40), ('deltanet', 26)])
def test_train_step_parity_at_reference_sizes(name, N):
    _check(name, N, seed=100 + N)


def test_bench_workload_forward_512():
    """adenet_3stream, 512 utterances x 40 frames (the bench batch): probabilities and argmax against the oracle."""
    _check('adenet_3stream', 512, seed=7, train=False)


def test_bench_workload_padded_layout():
    """The same network in the reference's padded layout (packed execution off) holds the same gates."""
    _check('adenet_3stream', 64, seed=164, packed='off')
