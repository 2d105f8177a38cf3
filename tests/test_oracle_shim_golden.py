"""The oracle (and, under -m gpu, the CUDA kernels themselves) against vectors produced by EXECUTING the reference's own
`utils/signal.py`, `custom/objectives.py` and `custom/updates.py` with a NumPy stand-in for the handful of Theano / Lasagne
calls they make (tests/golden/make_theano_shim_golden.py -> tests/golden/theano_shim.npz).  Rows a2, a8, a9 of SURVEY 8."""
import os

import numpy as np
import pytest

from oracle import ops

SG = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'theano_shim.npz'))


@pytest.mark.parametrize('k', [int(v) for v in SG['delta_cases']])
def test_delta_layer_oracle_is_the_reference_arithmetic_bit_for_bit(k):
    x, want, theta = SG['delta_x_%d' % k], SG['delta_y_%d' % k], int(SG['delta_theta_%d' % k])
    np.testing.assert_array_equal(ops.delta_fwd(x, theta), want)
    if x.size <= 2000:
        np.testing.assert_array_equal(ops.delta_fwd_literal(x, theta), want)


@pytest.mark.parametrize('k', range(int(SG['tsl_cases'])))
def test_temporal_softmax_loss_oracle(k):
    probs, y, mask, want = SG['tsl_probs_%d' % k], SG['tsl_y_%d' % k], SG['tsl_mask_%d' % k], float(SG['tsl_loss_%d' % k])
    loss, _ = ops.temporal_softmax_loss(probs, y, mask)
    assert abs(float(loss) - want) <= 2e-6 * abs(want)
    loss64, _ = ops.temporal_softmax_loss(probs, y, mask, np.float64)
    assert abs(float(loss64) - want) <= 2e-6 * abs(want)


def _adam_case():
    n, steps = int(SG['adam_nparams']), int(SG['adam_steps'])
    p0 = [SG['adam_p0_%d' % i].copy() for i in range(n)]
    grads = [[SG['adam_g_%d_%d' % (s, i)] for i in range(n)] for s in range(steps)]
    want = [[SG['adam_p_%d_%d' % (s, i)] for i in range(n)] for s in range(steps)]
    return p0, grads, want, [float(v) for v in SG['adam_lrs']]


def test_adam_vlr_oracle_follows_the_reference_update_rule():
    p, grads, want, lrs = _adam_case()
    st = {'t': np.float32(0), 'm': [np.zeros_like(a) for a in p], 'v': [np.zeros_like(a) for a in p]}
    for s, gs in enumerate(grads):
        ops.adam_step(p, gs, st, lrs)
        for a, b in zip(p, want[s]):
            # the same float32 expression, evaluated in the same order: equal up to the last bit of the step
            np.testing.assert_allclose(a, b, rtol=0, atol=2e-7 * max(1.0, float(np.abs(b).max())))
    assert float(st['t']) == len(grads)


# ---- the CUDA kernels against the same vectors, through the C-ABI ------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize('k', [int(v) for v in SG['delta_cases']])
def test_delta_kernel_matches_the_reference_vectors_bit_for_bit(k):
    import gpu_util as G
    x, want, theta = SG['delta_x_%d' % k], SG['delta_y_%d' % k], int(SG['delta_theta_%d' % k])
    N, T, F = x.shape
    ldx, ldy = (F + 3) // 4 * 4, (3 * F + 3) // 4 * 4
    xp = np.zeros((N * T, ldx), 'float32')
    xp[:, :F] = x.reshape(N * T, F)
    dx, dy = G.dev(xp), G.zeros((N * T, ldy))
    G.call('ipavsr_delta_fwd', dx.data_ptr(), ldx, dy.data_ptr(), ldy, N, T, F, theta, 1, G.stream())
    np.testing.assert_array_equal(G.host(dy)[:, :3 * F].reshape(N, T, 3 * F), want)


@pytest.mark.gpu
@pytest.mark.parametrize('k', range(int(SG['tsl_cases'])))
def test_temporal_softmax_loss_kernel_matches_the_reference_vectors(k):
    import gpu_util as G
    probs, y, mask, want = SG['tsl_probs_%d' % k], SG['tsl_y_%d' % k], SG['tsl_mask_%d' % k], float(SG['tsl_loss_%d' % k])
    N, T, V = probs.shape
    ld = (V + 3) // 4 * 4
    pp = np.zeros((N * T, ld), 'float32')
    pp[:, :V] = probs.reshape(N * T, V)
    d_p, d_y, d_m = G.dev(pp), G.dev(y), G.dev(mask)
    ls, dl = G.zeros((4,)), G.zeros((N * T, ld))
    cnt = float(mask.sum())
    G.call('ipavsr_temporal_softmax_loss', d_p.data_ptr(), ld, d_y.data_ptr(), d_m.data_ptr(), ls.data_ptr(), dl.data_ptr(),
           ld, N * T, V, 1.0 / cnt, None, G.stream())
    assert abs(float(G.host(ls)[0]) / cnt - want) <= 1e-5 * abs(want)


@pytest.mark.gpu
def test_adam_vlr_kernel_matches_the_reference_vectors():
    """ipavsr_optim_step (adam, per-tensor learning rates over 256-float segments) through 4 steps of the reference's own
    adam_vlr on 4 tensors with 3 different rates."""
    import gpu_util as G
    p0, grads, want, lrs = _adam_case()
    SEG = 256
    offs, n = [], 0
    for a in p0:
        offs.append(n)
        n += (a.size + SEG - 1) // SEG * SEG
    flat = np.zeros(n, 'float32')
    seg_id = np.zeros(n // SEG, 'int32')
    for i, (a, o) in enumerate(zip(p0, offs)):
        flat[o:o + a.size] = a.ravel()
        seg_id[o // SEG:(o + a.size + SEG - 1) // SEG] = i
    d_p, s1, s2 = G.dev(flat), G.zeros((n,)), G.zeros((n,))
    d_sl, d_si = G.dev(np.array(lrs, 'float32')), G.dev(seg_id)
    t = np.float32(0)
    for s, gs in enumerate(grads):
        g = np.zeros(n, 'float32')
        for a, o in zip(gs, offs):
            g[o:o + a.size] = a.ravel()
        t = np.float32(t + 1)
        sc = float(np.sqrt(np.float32(1) - np.float32(0.999) ** t) / (np.float32(1) - np.float32(0.9) ** t))
        d_g = G.dev(g)
        G.call('ipavsr_optim_step', 0, d_p.data_ptr(), d_g.data_ptr(), s1.data_ptr(), s2.data_ptr(), n, 0.0, d_sl.data_ptr(),
               d_si.data_ptr(), sc, 0.9, 0.999, 1e-8, 1.0, G.stream())
        got = G.host(d_p)
        for a, o in zip(want[s], offs):
            np.testing.assert_allclose(got[o:o + a.size].reshape(a.shape), a, rtol=0,
                                       atol=1e-6 * max(1.0, float(np.abs(a).max())))
