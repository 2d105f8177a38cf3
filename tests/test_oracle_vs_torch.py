"""Validate the NumPy oracle (oracle/ops.py) against the independent torch-float64 autograd restatement and
against SURVEY Appendix C's derived known-answer vectors.  CPU only."""
import numpy as np
import pytest
import torch

from oracle import ops
import torch_ref as R



@pytest.fixture(autouse=True)
def _float64_default():
    """torch_ref builds float64 graphs; restore the default afterwards so that no other test module inherits it."""
    old = torch.get_default_dtype()
    torch.set_default_dtype(torch.float64)
    yield
    torch.set_default_dtype(old)


def _t(a, grad=True):
    return torch.tensor(np.asarray(a, np.float64), requires_grad=grad)


def test_delta_known_answers_appendix_c():
    seqs = np.array([[[1, 2, 3, 4, 5], [10, 12, 13, 14, 15], [300, 1, 23, 56, 22]],
                     [[1, 1, 1, 1, 1], [1, 1, 100, 1, 1], [1, 1, 1, 1, 1]]], dtype='float32')   # utils/signal.py:95-100
    out = ops.delta_fwd(seqs, 1)
    exp0 = [[1, 2, 3, 4, 5, 4.5, 5, 5, 5, 5, 72.5, -2.75, 2.5, 10.5, 1.75],
            [10, 12, 13, 14, 15, 149.5, -0.5, 10, 26, 8.5, 70.25, -5.25, 0, 8, -0.75],
            [300, 1, 23, 56, 22, 145, -5.5, 5, 21, 3.5, -2.25, -2.5, -2.5, -2.5, -2.5]]
    exp1 = [[1, 1, 1, 1, 1, 0, 0, 49.5, 0, 0, 0, 0, -24.75, 0, 0],
            [1, 1, 100, 1, 1, 0, 0, 0, 0, 0, 0, 0, -49.5, 0, 0],
            [1, 1, 1, 1, 1, 0, 0, -49.5, 0, 0, 0, 0, -24.75, 0, 0]]
    np.testing.assert_array_equal(out[0], np.array(exp0, 'float32'))
    np.testing.assert_array_equal(out[1], np.array(exp1, 'float32'))


@pytest.mark.parametrize('theta', [1, 4, 9])
def test_delta_fwd_literal_and_torch(theta):
    rng = np.random.default_rng(theta)
    x = rng.normal(size=(3, 11, 7)).astype('float32')
    a = ops.delta_fwd(x, theta)
    b = ops.delta_fwd_literal(x, theta)
    np.testing.assert_array_equal(a, b)                       # vectorised == literal scan order, bitwise
    c = R.delta_layer(_t(x, False), theta).numpy()
    np.testing.assert_allclose(a, c, rtol=2e-6, atol=2e-6)
    # Theta=1 is the central difference (SURVEY 8c invariant 4)
    if theta == 1:
        np.testing.assert_allclose(a[:, 1:-1, 7:14], (x[:, 2:] - x[:, :-2]) / 2, rtol=1e-6, atol=1e-6)


@pytest.mark.parametrize('theta', [1, 3, 9])
def test_delta_bwd(theta):
    rng = np.random.default_rng(10 + theta)
    x = rng.normal(size=(2, 13, 5))
    g = rng.normal(size=(2, 13, 15))
    xt = _t(x)
    (R.delta_layer(xt, theta) * _t(g, False)).sum().backward()
    np.testing.assert_allclose(ops.delta_bwd(g, theta, np.float64), xt.grad.numpy(), rtol=1e-10, atol=1e-10)


@pytest.mark.parametrize('act', ['linear', 'sigmoid', 'rectify', 'tanh', 'leaky_rectify', 'softplus', 'elu'])
def test_dense(act):
    rng = np.random.default_rng(3)
    x, W, b, dy = rng.normal(size=(9, 6)), rng.normal(size=(6, 4)), rng.normal(size=(4,)), rng.normal(size=(9, 4))
    code = ops.act_code(act)
    y, cache = ops.dense_fwd(x, W, b, code, np.float64)
    dx, dW, db = ops.dense_bwd(dy, cache, W, code, np.float64)
    xt, Wt, bt = _t(x), _t(W), _t(b)
    z = xt @ Wt + bt
    f = {'linear': lambda v: v, 'sigmoid': torch.sigmoid, 'rectify': torch.relu, 'tanh': torch.tanh,
         'leaky_rectify': lambda v: torch.nn.functional.leaky_relu(v, 0.01),
         'softplus': torch.nn.functional.softplus, 'elu': torch.nn.functional.elu}[act]
    yt = f(z)
    (yt * _t(dy, False)).sum().backward()
    np.testing.assert_allclose(y, yt.detach().numpy(), rtol=1e-12, atol=1e-12)
    for a, t in ((dx, xt), (dW, Wt), (db, bt)):
        np.testing.assert_allclose(a, t.grad.numpy(), rtol=1e-10, atol=1e-10)


def _lstm_case(rng, N, T, I, H, peep, lens):
    p = {'W_in': rng.normal(0, .4, (I, 4 * H)), 'W_hid': rng.normal(0, .4, (H, 4 * H)), 'b': rng.normal(0, .2, (4 * H,)),
         'cell_init': rng.normal(0, .3, (H,)), 'hid_init': rng.normal(0, .3, (H,))}
    if peep:
        p['peep'] = rng.normal(0, .3, (3, H))
    x = rng.normal(size=(N, T, I))
    mask = (np.arange(T)[None, :] < np.asarray(lens)[:, None]).astype('uint8')
    return p, x, mask


@pytest.mark.parametrize('backwards', [False, True])
@pytest.mark.parametrize('peep', [False, True])
@pytest.mark.parametrize('scale', [1.0, 40.0])      # 40x output-gradient => the +-5 gate-gradient clip is active
def test_lstm_fwd_bwd(backwards, peep, scale):
    rng = np.random.default_rng(7)
    N, T, I, H = 4, 7, 5, 6
    p, x, mask = _lstm_case(rng, N, T, I, H, peep, [7, 3, 1, 5])
    dout = rng.normal(size=(N, T, H)) * scale
    out, cache = ops.lstm_fwd(x, mask, p, backwards, np.float64)
    dx, gr = ops.lstm_bwd(dout, cache, 5.0, np.float64)
    tp = {k: _t(v) for k, v in p.items()}
    xt = _t(x)
    ot = R.lstm(xt, torch.tensor(mask), tp['W_in'], tp['W_hid'], tp['b'], tp.get('peep'), tp['cell_init'],
                tp['hid_init'], backwards, 5.0)
    (ot * _t(dout, False)).sum().backward()
    np.testing.assert_allclose(out, ot.detach().numpy(), rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(dx, xt.grad.numpy(), rtol=1e-9, atol=1e-9)
    for k in gr:
        np.testing.assert_allclose(gr[k], tp[k].grad.numpy(), rtol=1e-9, atol=1e-9, err_msg=k)
    if scale > 1:   # make sure the clip really was exercised
        ot2 = R.lstm(_t(x), torch.tensor(mask), tp['W_in'].detach(), tp['W_hid'].detach(), tp['b'].detach(),
                     None if not peep else tp['peep'].detach(), tp['cell_init'].detach(),
                     tp['hid_init'].detach(), backwards, 0.0)
        assert ot2 is not None


def test_lstm_mask_semantics():
    """A.3: forward LSTM holds h[len-1] after len; backward LSTM outputs hid_init at t>=len; mask of ones ==
    unmasked; valid outputs do not depend on T."""
    rng = np.random.default_rng(11)
    N, T, I, H = 3, 8, 4, 5
    p, x, mask = _lstm_case(rng, N, T, I, H, True, [8, 5, 2])
    of, _ = ops.lstm_fwd(x, mask, p, False, np.float64)
    ob, _ = ops.lstm_fwd(x, mask, p, True, np.float64)
    np.testing.assert_array_equal(of[1, 5:], np.repeat(of[1, 4:5], 3, 0))
    np.testing.assert_array_equal(ob[2, 2:], np.repeat(p['hid_init'][None], 6, 0))
    o_short, _ = ops.lstm_fwd(x[1:2, :5], mask[1:2, :5], p, True, np.float64)
    np.testing.assert_allclose(ob[1, :5], o_short[0], rtol=1e-13)


def test_temporal_softmax_loss_and_grad():
    rng = np.random.default_rng(5)
    N, T, C = 3, 6, 4
    z = rng.normal(size=(N * T, C))
    probs = ops.softmax_rows(z).reshape(N, T, C)
    y = rng.integers(0, C, size=(N, T))
    mask = (np.arange(T)[None] < np.array([6, 2, 4])[:, None]).astype('uint8')
    loss, dp = ops.temporal_softmax_loss(probs, y, mask, np.float64)
    pt = _t(probs)
    lt = R.temporal_softmax_loss(pt, torch.tensor(y), torch.tensor(mask))
    lt.backward()
    np.testing.assert_allclose(loss, lt.item(), rtol=1e-12)
    np.testing.assert_allclose(dp, pt.grad.numpy(), rtol=1e-10, atol=1e-12)


def test_categorical_crossentropy():
    rng = np.random.default_rng(6)
    probs = ops.softmax_rows(rng.normal(size=(5, 4)))
    y = rng.integers(0, 4, size=(5,))
    loss, dp = ops.categorical_crossentropy_mean(probs, y, np.float64)
    pt = _t(probs)
    lt = -torch.log(pt[torch.arange(5), torch.tensor(y)]).mean()
    lt.backward()
    np.testing.assert_allclose(loss, lt.item(), rtol=1e-12)
    np.testing.assert_allclose(dp, pt.grad.numpy(), rtol=1e-10)


def test_batchnorm():
    rng = np.random.default_rng(8)
    x, beta, gamma, dy = rng.normal(size=(12, 5)), rng.normal(size=5), rng.normal(size=5), rng.normal(size=(12, 5))
    y, cache, new = ops.bn_fwd(x, beta, gamma, np.zeros(5), np.ones(5), False, 1e-4, 0.1, np.float64)
    dx, dbeta, dgamma = ops.bn_bwd(dy, cache, gamma, np.float64)
    xt, bt, gt = _t(x), _t(beta), _t(gamma)
    m = xt.mean(0)
    istd = 1 / torch.sqrt(xt.var(0, unbiased=False) + 1e-4)
    yt = (xt - m) * (gt * istd) + bt
    (yt * _t(dy, False)).sum().backward()
    np.testing.assert_allclose(y, yt.detach().numpy(), rtol=1e-12)
    np.testing.assert_allclose(dx, xt.grad.numpy(), rtol=1e-9, atol=1e-11)
    np.testing.assert_allclose(dbeta, bt.grad.numpy(), rtol=1e-10)
    np.testing.assert_allclose(dgamma, gt.grad.numpy(), rtol=1e-10)
    np.testing.assert_allclose(new[0], 0.1 * x.mean(0), rtol=1e-12)
    np.testing.assert_allclose(new[1], 0.9 + 0.1 * istd.detach().numpy(), rtol=1e-12)


def test_adam_matches_torch():
    rng = np.random.default_rng(9)
    p0 = rng.normal(size=(7, 3)).astype('float32')
    gs = [rng.normal(size=(7, 3)).astype('float32') for _ in range(5)]
    p = p0.copy()
    st = {'t': np.float32(0), 'm': [np.zeros_like(p)], 'v': [np.zeros_like(p)]}
    for g in gs:
        ops.adam_step([p], [g], st, [1e-3])
    pt = torch.tensor(p0, dtype=torch.float32, requires_grad=True)
    opt = torch.optim.Adam([pt], lr=1e-3, betas=(0.9, 0.999), eps=1e-8)
    for g in gs:
        pt.grad = torch.tensor(g)
        opt.step()
    # Lasagne's epsilon sits outside the bias correction (eps vs eps*sqrt(1-b2^t)) — tiny, documented difference
    np.testing.assert_allclose(p, pt.detach().numpy(), rtol=0, atol=2e-6)
