"""CPU ORACLE — test infrastructure only.  NOT part of the product path.

NumPy restatement of every arithmetic op on the AdeNet/DeltaNet hot path, forward and
backward.  Only `tests/`, `__graft_entry__.smoke()` and the `cpu_baseline` / `--impl reference` legs of
`bench.py` may import this package; `ipavsr_b200/` never does.

PARITY STATUS: *partially pinned*.
  * `oracle/preprocessing.py` (a10–a14) is pinned against outputs of the reference's own
    `utils/preprocessing.py` executed in the build container (`tests/golden/make_golden.py`).
  * The parts of the Theano-level arithmetic that live in the REFERENCE's own sources — the DeltaLayer
    (`utils/signal.py:7-80`, a2), `temporal_softmax_loss` (`custom/objectives.py:27-37`, a8) and `adam_vlr` /
    `generate_lr_map` (`custom/updates.py:10-99`, a9) — are pinned against vectors made by executing those source files,
    unmodified, with a NumPy stand-in for the few Theano / Lasagne calls they make
    (`tests/golden/make_theano_shim_golden.py` -> `tests/golden/theano_shim.npz`, checked in
    `tests/test_oracle_shim_golden.py`: `delta_fwd` bit for bit).  That is the reference's code run statement by statement,
    not its compiled Theano graph.
  * What lives inside Lasagne / nolearn (Dense, LSTMLayer, BatchNorm, Dropout, the other update rules and objectives;
    rows a1, a3–a7, f4) is **parity unpinned**: Theano, Lasagne and nolearn are not in `/root/reference`, are pinned by the
    reference only as "master" (`README.md:30-33`), are not installable here, and the reference ships no golden vectors
    for them.  These functions restate the published Lasagne semantics; they are
    cross-checked against an independent torch-float64 autograd restatement and finite differences in
    `tests/test_oracle_*.py`, and against the hand-derived vectors of SURVEY Appendix C.

All functions take `dt` (np.float32 reproduces the reference's floatX=float32 arithmetic with BLAS dots;
np.float64 gives a high-precision value used to separate kernel error from float32 rounding of the oracle).
"""
import numpy as np

# activation codes shared with include/ipavsr_b200.h (IPAVSR_ACT_*)
ACT_LINEAR, ACT_SIGMOID, ACT_RECTIFY, ACT_TANH, ACT_LEAKY, ACT_VERY_LEAKY, ACT_SOFTPLUS, ACT_ELU, ACT_SOFTMAX = range(9)
_ACT_BY_NAME = {'linear': 0, 'sigmoid': 1, 'rectify': 2, 'tanh': 3, 'leaky_rectify': 4,
                'very_leaky_rectify': 5, 'softplus': 6, 'elu': 7, 'softmax': 8}


def act_code(nl):
    return _ACT_BY_NAME[nl if isinstance(nl, str) else nl.name]


def sigmoid(x):
    return 1.0 / (1.0 + np.exp(-x))


def softmax_rows(z):
    e = np.exp(z - z.max(axis=1, keepdims=True))
    return e / e.sum(axis=1, keepdims=True)


# ---------------------------------------------------------------------------------------------------
# a1  DenseLayer (modelzoo/pretrained_encoder.py:4-9)
# ---------------------------------------------------------------------------------------------------
def act_fwd(z, act):
    if act == ACT_LINEAR:
        return z
    if act == ACT_SIGMOID:
        return sigmoid(z)
    if act == ACT_RECTIFY:
        return 0.5 * (z + np.abs(z))            # Lasagne rectify
    if act == ACT_TANH:
        return np.tanh(z)
    if act in (ACT_LEAKY, ACT_VERY_LEAKY):
        a = 0.01 if act == ACT_LEAKY else 1.0 / 3.0
        return (0.5 * (1 + a)) * z + (0.5 * (1 - a)) * np.abs(z)
    if act == ACT_SOFTPLUS:
        return np.logaddexp(0, z)
    if act == ACT_ELU:
        return np.where(z > 0, z, np.exp(np.minimum(z, 0)) - 1)
    if act == ACT_SOFTMAX:
        return softmax_rows(z)
    raise ValueError(act)


def act_bwd(dy, z, y, act):
    """dL/dz from dL/dy.  Rectify keeps Theano's sub-gradient 0.5 at exactly z == 0."""
    if act == ACT_LINEAR:
        return dy
    if act == ACT_SIGMOID:
        return dy * y * (1 - y)
    if act == ACT_RECTIFY:
        return dy * (0.5 * (1 + np.sign(z)))
    if act == ACT_TANH:
        return dy * (1 - y * y)
    if act in (ACT_LEAKY, ACT_VERY_LEAKY):
        a = 0.01 if act == ACT_LEAKY else 1.0 / 3.0
        return dy * ((0.5 * (1 + a)) + (0.5 * (1 - a)) * np.sign(z))
    if act == ACT_SOFTPLUS:
        return dy * sigmoid(z)
    if act == ACT_ELU:
        return dy * np.where(z > 0, 1.0, y + 1.0)
    if act == ACT_SOFTMAX:
        return y * (dy - (dy * y).sum(axis=1, keepdims=True))
    raise ValueError(act)


def dense_fwd(x, W, b, act, dt=np.float32):
    z = np.dot(x.astype(dt), W.astype(dt))
    if b is not None:
        z = z + b.astype(dt)
    y = act_fwd(z, act).astype(dt)
    return y, (x.astype(dt), z, y)


def dense_bwd(dy, cache, W, act, dt=np.float32, need_dx=True):
    x, z, y = cache
    dz = act_bwd(dy.astype(dt), z, y, act).astype(dt)
    dW = np.dot(x.T, dz)
    db = dz.sum(axis=0)
    dx = np.dot(dz, W.astype(dt).T) if need_dx else None
    return dx, dW.astype(dt), db.astype(dt)


# ---------------------------------------------------------------------------------------------------
# a2  DeltaLayer (custom/layers.py:105-121 -> utils/signal.py:59-80 -> :26-39 -> :7-23)
# ---------------------------------------------------------------------------------------------------
def _delta_coeff(A, theta):
    """utils/signal.py delta_coeff: A is (N,T,F) float32; returns float32.

    Per theta the reference evaluates `theta * (Y[t+theta] - Y[t-theta]) / (2*theta*theta)` with an int32
    theta, which Theano upcasts to float64, adds it to the float32 accumulator and rounds the sum back to
    float32 (`utils/signal.py:19-21`).  Edge handling: theta replicated first/last frames (:68-69)."""
    N, T, F = A.shape
    t = np.arange(T)
    d = np.zeros((N, T, F), dtype=np.float32)
    for th in range(1, theta + 1):
        hi = A[:, np.minimum(t + th, T - 1), :]
        lo = A[:, np.maximum(t - th, 0), :]
        diff = (hi - lo).astype(np.float32)                         # float32 subtraction
        term = np.float64(th) * diff.astype(np.float64) / np.float64(2 * th * th)
        d = (d.astype(np.float64) + term).astype(np.float32)
    return d


def delta_fwd(x, theta):
    """append_delta_coeff over a batch: (N,T,F) float32 -> (N,T,3F) float32; mask-agnostic, whole padded T."""
    x = np.ascontiguousarray(x, dtype=np.float32)
    d = _delta_coeff(x, theta)
    a = _delta_coeff(d, theta)
    return np.concatenate([x, d, a], axis=2)


def delta_fwd_literal(x, theta):
    """The same arithmetic as the three nested theano.scan loops (n, t, theta) — the reference's cost shape.
    Pure Python; only for small cases and for the `cpu_baseline` cost-shape figure."""
    x = np.asarray(x, dtype=np.float32)
    N, T, F = x.shape

    def coeff(A):
        Y = np.concatenate([np.repeat(A[:1], theta, 0), A, np.repeat(A[-1:], theta, 0)], 0)
        out = np.zeros_like(A)
        for t in range(T):
            acc = np.zeros((F,), np.float32)
            for th in range(1, theta + 1):
                dth = np.float64(th) * np.float64(1) * (Y[theta + t + th] - Y[theta + t - th]).astype(np.float64) \
                    / np.float64(2 * th * th)
                acc = (acc.astype(np.float64) + dth).astype(np.float32)
            out[t] = acc
        return out

    res = np.zeros((N, T, 3 * F), np.float32)
    for n in range(N):
        d = coeff(x[n])
        a = coeff(d)
        res[n] = np.concatenate([x[n], d, a], 1)
    return res


def delta_matrix(T, theta):
    """The banded T x T operator D with d = D x (float64), clamp folded in (SURVEY A.2 closed form)."""
    D = np.zeros((T, T), dtype=np.float64)
    for t in range(T):
        for th in range(1, theta + 1):
            D[t, min(t + th, T - 1)] += 1.0 / (2 * th)
            D[t, max(t - th, 0)] -= 1.0 / (2 * th)
    return D


def delta_bwd(g, theta, dt=np.float32):
    """Gradient w.r.t. the (N,T,F) input given g (N,T,3F):  gx + D^T (gd + D^T ga)."""
    N, T, F3 = g.shape
    F = F3 // 3
    D = delta_matrix(T, theta)
    g = g.astype(np.float64)
    gx, gd, ga = g[:, :, :F], g[:, :, F:2 * F], g[:, :, 2 * F:]
    inner = gd + np.einsum('ts,ntf->nsf', D, ga)
    out = gx + np.einsum('ts,ntf->nsf', D, inner)
    return out.astype(dt)


# ---------------------------------------------------------------------------------------------------
# a3  Lasagne LSTMLayer (custom/layers.py:10-80)
# ---------------------------------------------------------------------------------------------------
def lstm_fwd(x, mask, p, backwards=False, dt=np.float32):
    """x (N,T,I), mask (N,T) {0,1}.  p: dict with stacked W_in (I,4H), W_hid (H,4H), b (4H,), optional
    peep (3,H) rows (ci, cf, co), cell_init (H,), hid_init (H,).  Gate order [i|f|c|o].  Returns out (N,T,H)
    and the cache for lstm_bwd."""
    x = x.astype(dt)
    N, T, I = x.shape
    W_in, W_hid, b = p['W_in'].astype(dt), p['W_hid'].astype(dt), p['b'].astype(dt)
    H = W_hid.shape[0]
    peep = p.get('peep')
    if peep is not None:
        peep = peep.astype(dt)
    xW = (np.dot(x.reshape(N * T, I), W_in) + b).reshape(N, T, 4 * H)
    c_prev = np.repeat(p['cell_init'].astype(dt).reshape(1, H), N, 0)
    h_prev = np.repeat(p['hid_init'].astype(dt).reshape(1, H), N, 0)
    out = np.zeros((N, T, H), dt)
    steps = []
    order = range(T - 1, -1, -1) if backwards else range(T)
    for t in order:
        g = xW[:, t] + np.dot(h_prev, W_hid)
        gi, gf, gc, go = g[:, :H], g[:, H:2 * H], g[:, 2 * H:3 * H], g[:, 3 * H:]
        if peep is not None:
            gi = gi + c_prev * peep[0]
            gf = gf + c_prev * peep[1]
        i, f, cin = sigmoid(gi), sigmoid(gf), np.tanh(gc)
        c_u = f * c_prev + i * cin
        if peep is not None:
            go = go + c_u * peep[2]
        o = sigmoid(go)
        tc = np.tanh(c_u)
        h_u = o * tc
        m = mask[:, t].astype(bool)[:, None]
        c = np.where(m, c_u, c_prev).astype(dt)
        h = np.where(m, h_u, h_prev).astype(dt)
        steps.append((t, i, f, cin, o, tc, c_u, c_prev, h_prev, m))
        out[:, t] = h
        c_prev, h_prev = c, h
    return out, (x, steps, p, backwards)


def lstm_bwd(dout, cache, clip=5.0, dt=np.float32, need_dx=True):
    x, steps, p, backwards = cache
    N, T, I = x.shape
    W_in, W_hid = p['W_in'].astype(dt), p['W_hid'].astype(dt)
    H = W_hid.shape[0]
    peep = p.get('peep')
    if peep is not None:
        peep = peep.astype(dt)
    dpeep = np.zeros((3, H), dt) if peep is not None else None
    dxW = np.zeros((N, T, 4 * H), dt)
    dW_hid = np.zeros_like(W_hid)
    dh_next = np.zeros((N, H), dt)
    dc_next = np.zeros((N, H), dt)
    dout = dout.astype(dt)
    for (t, i, f, cin, o, tc, c_u, c_prev, h_prev, m) in reversed(steps):
        dh = dout[:, t] + dh_next
        dc = dc_next
        dh_pass, dc_pass = np.where(m, 0, dh), np.where(m, 0, dc)
        dh_u, dc_u = np.where(m, dh, 0), np.where(m, dc, 0)
        dgo = dh_u * tc * o * (1 - o)
        dc_u = dc_u + dh_u * o * (1 - tc * tc)
        if peep is not None:
            dc_u = dc_u + dgo * peep[2]
            dpeep[2] += (dgo * c_u).sum(0)
        dgi = dc_u * cin * i * (1 - i)
        dgf = dc_u * c_prev * f * (1 - f)
        dgc = dc_u * i * (1 - cin * cin)
        dc_prev = dc_u * f
        if peep is not None:
            dc_prev = dc_prev + dgi * peep[0] + dgf * peep[1]
            dpeep[0] += (dgi * c_prev).sum(0)
            dpeep[1] += (dgf * c_prev).sum(0)
        dg = np.concatenate([dgi, dgf, dgc, dgo], axis=1)
        if clip:
            dg = np.clip(dg, -clip, clip)          # theano.gradient.grad_clip on the pre-peephole gates
        dg = dg.astype(dt)
        dxW[:, t] = dg
        dW_hid += np.dot(h_prev.T, dg)
        dh_next = (np.dot(dg, W_hid.T) + dh_pass).astype(dt)
        dc_next = (dc_prev + dc_pass).astype(dt)
    grads = {'W_in': np.dot(x.reshape(N * T, I).T, dxW.reshape(N * T, 4 * H)).astype(dt),
             'W_hid': dW_hid, 'b': dxW.reshape(N * T, 4 * H).sum(0).astype(dt),
             'cell_init': dc_next.sum(0).astype(dt), 'hid_init': dh_next.sum(0).astype(dt)}
    if peep is not None:
        grads['peep'] = dpeep
    dx = np.dot(dxW.reshape(N * T, 4 * H), W_in.T).reshape(N, T, I).astype(dt) if need_dx else None
    return dx, grads


# ---------------------------------------------------------------------------------------------------
# a6  BatchNormLayer (modelzoo/adenet_v1.py:82)
# ---------------------------------------------------------------------------------------------------
def bn_fwd(x, beta, gamma, mean, inv_std, deterministic, eps=1e-4, alpha=0.1, dt=np.float32):
    x = x.astype(dt)
    if deterministic:
        y = (x - mean.astype(dt)) * (gamma.astype(dt) * inv_std.astype(dt)) + beta.astype(dt)
        return y.astype(dt), None, (mean, inv_std)
    mb = x.mean(axis=0)
    ib = 1.0 / np.sqrt(x.var(axis=0) + dt(eps))
    y = (x - mb) * (gamma.astype(dt) * ib) + beta.astype(dt)
    new_mean = (1 - dt(alpha)) * mean.astype(dt) + dt(alpha) * mb
    new_inv_std = (1 - dt(alpha)) * inv_std.astype(dt) + dt(alpha) * ib
    return y.astype(dt), (x, mb.astype(dt), ib.astype(dt)), (new_mean.astype(dt), new_inv_std.astype(dt))


def bn_bwd(dy, cache, gamma, dt=np.float32):
    x, mb, ib = cache
    dy = dy.astype(dt)
    M = x.shape[0]
    xh = (x - mb) * ib
    dbeta = dy.sum(0)
    dgamma = (dy * xh).sum(0)
    dx = (gamma.astype(dt) * ib / M) * (M * dy - dbeta - xh * dgamma)
    return dx.astype(dt), dbeta.astype(dt), dgamma.astype(dt)


# ---------------------------------------------------------------------------------------------------
# a8  losses (custom/objectives.py:4-39; avletters/trimodal.py:327)
# ---------------------------------------------------------------------------------------------------
def temporal_softmax_loss(probs, y, mask, dt=np.float32):
    """probs (N,T,C) are already softmax outputs; the reference softmaxes them *again* (:34-35).
    Returns loss and dL/dprobs."""
    N, T, C = probs.shape
    x = probs.reshape(N * T, C).astype(dt)
    yf = y.reshape(N * T).astype(np.int64)
    mf = mask.reshape(N * T).astype(dt)
    total = mf.sum()
    q = softmax_rows(x)
    loss = -(mf * np.log(q[np.arange(N * T), yf])).sum() / total
    dq = q.copy()
    dq[np.arange(N * T), yf] -= 1
    dprobs = dq * (mf / total)[:, None]
    return dt(loss), dprobs.reshape(N, T, C).astype(dt)


def categorical_crossentropy_mean(probs, y, dt=np.float32):
    """T.mean(categorical_crossentropy(pred, y)): probs (N,C), y (N,) int."""
    N = probs.shape[0]
    p = probs.astype(dt)
    yy = y.astype(np.int64)
    loss = -np.log(p[np.arange(N), yy]).mean()
    dp = np.zeros_like(p)
    dp[np.arange(N), yy] = -1.0 / (p[np.arange(N), yy] * N)
    return dt(loss), dp


def squared_error_mean(pred, target, dt=np.float32):
    """T.mean(lasagne.objectives.squared_error(pred, target)) — nolearn's regression objective
    (`avletters/trimodal.py:81`): mean over every element; returns (loss, d loss / d pred)."""
    p = np.asarray(pred, dt)
    d = p - np.asarray(target, dt).reshape(p.shape)
    return dt((d * d).mean()), (2.0 / d.size) * d


# ---------------------------------------------------------------------------------------------------
# a9  update rules (custom/updates.py:35-99; Lasagne adam/adadelta/sgd/momentum)
# ---------------------------------------------------------------------------------------------------
def adam_step(params, grads, state, lrs, beta1=0.9, beta2=0.999, eps=1e-8):
    """In-place float32 Adam with a per-parameter learning rate list (adam_vlr; plain adam = equal lrs).
    state: {'t': float32 scalar, 'm': [...], 'v': [...]}."""
    one = np.float32(1)
    t = np.float32(state['t'] + one)
    b1, b2 = np.float32(beta1), np.float32(beta2)
    for k, (p, g) in enumerate(zip(params, grads)):
        a_t = np.float32(lrs[k]) * np.sqrt(one - b2 ** t) / (one - b1 ** t)
        state['m'][k] = b1 * state['m'][k] + (one - b1) * g
        state['v'][k] = b2 * state['v'][k] + (one - b2) * g ** 2
        p -= (a_t * state['m'][k] / (np.sqrt(state['v'][k]) + np.float32(eps))).astype(np.float32)
    state['t'] = t


def adadelta_step(params, grads, state, lr=1.0, rho=0.95, eps=1e-6):
    one = np.float32(1)
    rho, eps, lr = np.float32(rho), np.float32(eps), np.float32(lr)
    for k, (p, g) in enumerate(zip(params, grads)):
        acc = rho * state['acc'][k] + (one - rho) * g ** 2
        upd = g * np.sqrt(state['dacc'][k] + eps) / np.sqrt(acc + eps)
        p -= lr * upd
        state['dacc'][k] = rho * state['dacc'][k] + (one - rho) * upd ** 2
        state['acc'][k] = acc


def sgd_momentum_step(params, grads, state, lr, momentum=0.9, nesterov=False):
    lr, mu = np.float32(lr), np.float32(momentum)
    for k, (p, g) in enumerate(zip(params, grads)):
        v = mu * state['vel'][k] - lr * g
        state['vel'][k] = v
        if nesterov:
            p += mu * v - lr * g
        else:
            p += v


def sgd_step(params, grads, lr):
    for p, g in zip(params, grads):
        p -= np.float32(lr) * g
