"""Parity of every CUDA kernel against the NumPy oracle, through the C-ABI (run on the B200 with -m gpu)."""
import os

import numpy as np
import pytest
import torch

from oracle import ops, preprocessing as OP
import gpu_util as G

pytestmark = pytest.mark.gpu

GOLD = np.load(os.path.join(os.path.dirname(__file__), 'golden', 'preprocessing.npz'))


# ---------------------------------------------------------------------------------------------------------
# GEMM (FP32 CUDA-core mode): all transposes, odd shapes, bias/activation/accumulate, split-K
# ---------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize('M,N,K', [(1040, 2000, 1200), (129, 50, 500), (26, 26, 250), (50, 500, 20480),
                                   (1, 7, 3), (300, 1000, 150), (64, 64, 16), (257, 130, 91)])
@pytest.mark.parametrize('ta,tb', [(0, 0), (0, 1), (1, 0), (1, 1)])
def test_gemm_fp32(M, N, K, ta, tb):
    rng = np.random.default_rng(M * 7 + N * 3 + K + ta * 2 + tb)
    A = rng.normal(size=(K, M) if ta else (M, K)).astype('float32')
    B = rng.normal(size=(N, K) if tb else (K, N)).astype('float32')
    bias = rng.normal(size=(N,)).astype('float32')
    C0 = rng.normal(size=(M, N)).astype('float32')
    ref = (A.T if ta else A).astype(np.float64) @ (B.T if tb else B).astype(np.float64)
    dA, dB, db = G.dev(A), G.dev(B), G.dev(bias)
    for act, acc, use_bias in [(0, 0, False), (1, 0, True), (2, 1, True), (0, 1, False)]:
        dC = G.dev(C0)
        G.call('ipavsr_gemm', 0, ta, tb, M, N, K, dA.data_ptr(), A.shape[1], dB.data_ptr(), B.shape[1],
               dC.data_ptr(), N, db.data_ptr() if use_bias else None, act, acc, None, 0, G.stream())
        z = ref + (C0 if acc else 0) + (bias if use_bias else 0)
        want = ops.act_fwd(z, act)
        got = G.host(dC)
        tol = 2e-6 * np.sqrt(K) + 1e-6
        assert G.relerr(got, want) < tol, (act, acc, G.relerr(got, want))


def test_gemm_strided_views():
    """leading dimensions larger than the logical width; a K-split accumulate over column segments."""
    rng = np.random.default_rng(5)
    M, K1, K2, N = 200, 250, 250, 1000
    X = rng.normal(size=(M, 504)).astype('float32')          # two 250-wide segments at ld 252 each side by side
    W = rng.normal(size=(K1 + K2, N)).astype('float32')
    dX, dW, dC = G.dev(X), G.dev(W), G.zeros((M, N))
    G.call('ipavsr_gemm', 0, 0, 0, M, N, K1, dX.data_ptr(), 504, dW.data_ptr(), N, dC.data_ptr(), N, None, 0, 0, None,
           0, G.stream())
    G.call('ipavsr_gemm', 0, 0, 0, M, N, K2, dX.data_ptr() + 4 * 252, 504, dW.data_ptr() + 4 * K1 * N, N,
           dC.data_ptr(), N, None, 0, 1, None, 0, G.stream())
    want = X[:, :250].astype(np.float64) @ W[:250] + X[:, 252:502].astype(np.float64) @ W[250:]
    assert G.relerr(G.host(dC), want) < 1e-5


# ---------------------------------------------------------------------------------------------------------
# DeltaLayer
# ---------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize('N,T,F,theta', [(2, 3, 5, 1), (26, 40, 50, 9), (7, 29, 90, 9), (5, 40, 30, 4), (3, 17, 50, 6),
                                         (1, 1, 4, 3), (300, 40, 50, 9), (4, 12, 51, 2), (1000, 40, 90, 9), (777, 29, 30, 4),
                                         (130, 48, 50, 1), (70, 60, 20, 9), (9, 5, 50, 9), (4, 2, 30, 9), (3, 1, 50, 9),
                                         (641, 40, 50, 9), (37, 24, 6, 4), (11, 25, 512, 9), (2, 40, 600, 9)])
def test_delta_fwd_exact(N, T, F, theta):
    rng = np.random.default_rng(N + T + F + theta)
    x = rng.normal(size=(N, T, F)).astype('float32')
    want = ops.delta_fwd(x, theta)
    ldx, ldy = (F + 3) // 4 * 4, (3 * F + 3) // 4 * 4
    xp = np.zeros((N * T, ldx), 'float32')
    xp[:, :F] = x.reshape(N * T, F)
    dx, dy = G.dev(xp), G.zeros((N * T, ldy))
    G.call('ipavsr_delta_fwd', dx.data_ptr(), ldx, dy.data_ptr(), ldy, N, T, F, theta, 1, G.stream())
    got = G.host(dy)[:, :3 * F].reshape(N, T, 3 * F)
    # exact mode reproduces the reference's float64 operation sequence: bit-identical
    np.testing.assert_array_equal(got, want)
    # fast (pure float32) mode: stated tolerance 1e-6 of the signal scale
    G.call('ipavsr_delta_fwd', dx.data_ptr(), ldx, dy.data_ptr(), ldy, N, T, F, theta, 0, G.stream())
    got2 = G.host(dy)[:, :3 * F].reshape(N, T, 3 * F)
    assert np.abs(got2 - want).max() <= 2e-6 * max(1.0, np.abs(want).max())


@pytest.mark.parametrize('T,theta', [(40, 9), (29, 4)])
def test_delta_wide_pitch_leaves_neighbours_untouched(T, theta):
    """Outputs that are column slices of a wider matrix (row pitch well beyond the data): the streaming kernels must take
    their register-store path and write nothing outside the [x | d | a] / g_x columns."""
    rng = np.random.default_rng(T + theta)
    N, F = 70, 50
    ldx, ldy = 64, 3 * F + 18
    x = rng.normal(size=(N, T, F)).astype('float32')
    xp = rng.normal(size=(N * T, ldx)).astype('float32')
    xp[:, :F] = x.reshape(N * T, F)
    ys = rng.normal(size=(N * T, ldy)).astype('float32')
    want = ops.delta_fwd(x, theta)
    for exact in (1, 0):
        dx, dy = G.dev(xp), G.dev(ys)
        G.call('ipavsr_delta_fwd', dx.data_ptr(), ldx, dy.data_ptr(), ldy, N, T, F, theta, exact, G.stream())
        got = G.host(dy)
        np.testing.assert_array_equal(got[:, 3 * F:], ys[:, 3 * F:])
        if exact:
            np.testing.assert_array_equal(got[:, :3 * F].reshape(N, T, 3 * F), want)
        else:
            assert np.abs(got[:, :3 * F].reshape(N, T, 3 * F) - want).max() <= 2e-6 * max(1.0, np.abs(want).max())
    g = rng.normal(size=(N, T, 3 * F)).astype('float32')
    gp = rng.normal(size=(N * T, ldy)).astype('float32')
    gp[:, :3 * F] = g.reshape(N * T, 3 * F)
    wantg = ops.delta_bwd(g, theta, np.float64)
    base = rng.normal(size=(N * T, ldx)).astype('float32')
    for acc in (0, 1):
        dg, dgx = G.dev(gp), G.dev(base)
        G.call('ipavsr_delta_bwd', dg.data_ptr(), ldy, dgx.data_ptr(), ldx, N, T, F, theta, acc, G.stream())
        got = G.host(dgx)
        np.testing.assert_array_equal(got[:, F:], base[:, F:])
        ref = wantg + (base[:, :F].reshape(N, T, F) if acc else 0)
        assert G.relerr(got[:, :F].reshape(N, T, F), ref) < 2e-6


def test_delta_known_answers():
    seqs = np.array([[[1, 2, 3, 4, 5], [10, 12, 13, 14, 15], [300, 1, 23, 56, 22]],
                     [[1, 1, 1, 1, 1], [1, 1, 100, 1, 1], [1, 1, 1, 1, 1]]], dtype='float32')
    xp = np.zeros((6, 8), 'float32')
    xp[:, :5] = seqs.reshape(6, 5)
    dx, dy = G.dev(xp), G.zeros((6, 16))
    G.call('ipavsr_delta_fwd', dx.data_ptr(), 8, dy.data_ptr(), 16, 2, 3, 5, 1, 1, G.stream())
    got = G.host(dy)[:, :15].reshape(2, 3, 15)
    np.testing.assert_array_equal(got[0, 0], [1, 2, 3, 4, 5, 4.5, 5, 5, 5, 5, 72.5, -2.75, 2.5, 10.5, 1.75])
    np.testing.assert_array_equal(got[1, 2], [1, 1, 1, 1, 1, 0, 0, -49.5, 0, 0, 0, 0, -24.75, 0, 0])


@pytest.mark.parametrize('N,T,F,theta', [(3, 11, 7, 2), (26, 40, 50, 9), (5, 1, 6, 4), (300, 40, 90, 9), (64, 29, 30, 4),
                                         (9, 5, 50, 9), (33, 2, 30, 1), (12, 48, 50, 1), (7, 24, 6, 9), (641, 40, 50, 9),
                                         (11, 25, 512, 9), (2, 40, 400, 9), (5, 60, 20, 9)])
def test_delta_bwd(N, T, F, theta):
    rng = np.random.default_rng(9)
    g = rng.normal(size=(N, T, 3 * F)).astype('float32')
    want = ops.delta_bwd(g, theta, np.float64)
    ldg, ldx = (3 * F + 3) // 4 * 4, (F + 3) // 4 * 4
    gp = np.zeros((N * T, ldg), 'float32')
    gp[:, :3 * F] = g.reshape(N * T, 3 * F)
    base = rng.normal(size=(N * T, ldx)).astype('float32')
    for acc in (0, 1):
        dg, dxx = G.dev(gp), G.dev(base)
        G.call('ipavsr_delta_bwd', dg.data_ptr(), ldg, dxx.data_ptr(), ldx, N, T, F, theta, acc, G.stream())
        got = G.host(dxx)[:, :F].reshape(N, T, F)
        ref = want + (base[:, :F].reshape(N, T, F) if acc else 0)
        assert G.relerr(got, ref) < 2e-6
    # linearity / adjointness: <D x, g> == <x, D^T g>
    x = rng.normal(size=(N, T, F)).astype('float32')
    lhs = (ops.delta_fwd(x, theta).astype(np.float64) * g).sum()
    rhs = (x.astype(np.float64) * want).sum()
    assert abs(lhs - rhs) < 1e-4 * max(1.0, abs(lhs))


# ---------------------------------------------------------------------------------------------------------
# LSTM recurrence, both implementations
# ---------------------------------------------------------------------------------------------------------
def _lstm_inputs(rng, N, T, I, H, peep, lens):
    p = {'W_in': rng.normal(0, .3, (I, 4 * H)).astype('float32'), 'W_hid': rng.normal(0, .3, (H, 4 * H)).astype('float32'),
         'b': rng.normal(0, .2, (4 * H,)).astype('float32'), 'cell_init': rng.normal(0, .3, (H,)).astype('float32'),
         'hid_init': rng.normal(0, .3, (H,)).astype('float32')}
    if peep:
        p['peep'] = rng.normal(0, .3, (3, H)).astype('float32')
    x = rng.normal(size=(N, T, I)).astype('float32')
    mask = (np.arange(T)[None, :] < np.asarray(lens)[:, None]).astype('uint8')
    return p, x, mask


@pytest.mark.parametrize('impl', [0, 1])
@pytest.mark.parametrize('N,T,I,H,peep,backwards,scale',
                         [(4, 7, 5, 6, True, False, 1.0), (4, 7, 5, 6, False, True, 40.0),
                          (26, 40, 150, 250, True, False, 1.0), (26, 40, 150, 250, False, True, 30.0),
                          (37, 12, 20, 70, True, True, 1.0), (33, 5, 8, 500, True, False, 1.0),
                          (3, 1, 4, 33, True, True, 1.0)])
def test_lstm_fwd_bwd(impl, N, T, I, H, peep, backwards, scale):
    rng = np.random.default_rng(N * 3 + T + H + impl)
    lens = rng.integers(1, T + 1, size=N)
    lens[0] = T
    p, x, mask = _lstm_inputs(rng, N, T, I, H, peep, lens)
    dout = (rng.normal(size=(N, T, H)) * scale).astype('float32')
    out_ref, cache = ops.lstm_fwd(x, mask, p, backwards, np.float64)
    dx_ref, gr = ops.lstm_bwd(dout, cache, 5.0, np.float64)
    dxw_ref = None
    # device inputs: xw precomputed (the hoisted projection is a separate GEMM, tested above)
    xw = (x.reshape(N * T, I).astype(np.float64) @ p['W_in'].astype(np.float64) + p['b']).astype('float32')
    ldh = (H + 3) // 4 * 4
    d_xw = G.dev(G.interleave_gates(xw, H))
    d_whid = G.dev(G.interleave_gates(p['W_hid'], H))
    d_peep = G.dev(p['peep']) if peep else None
    d_ci, d_hi, d_mask = G.dev(p['cell_init']), G.dev(p['hid_init']), G.dev(mask)
    d_out, d_gates = G.zeros((N * T, ldh)), G.zeros((N * T, 4 * H))
    d_cell, d_hprev = G.zeros((N * T, H)), G.zeros((N * T, ldh))
    nbytes = G.lib().ipavsr_lstm_workspace_bytes(N, T, H)
    ws = G.zeros(((nbytes + 3) // 4,))
    G.call('ipavsr_lstm_fwd', d_xw.data_ptr(), d_whid.data_ptr(), G.ptr(d_peep), d_ci.data_ptr(), d_hi.data_ptr(),
           d_mask.data_ptr(), d_out.data_ptr(), d_gates.data_ptr(), d_cell.data_ptr(), d_hprev.data_ptr(), N, T, H, ldh,
           int(backwards), impl, ws.data_ptr(), nbytes, G.stream())
    out = G.host(d_out)[:, :H].reshape(N, T, H)
    assert G.relerr(out, out_ref) < 2e-5, G.relerr(out, out_ref)
    # backward
    dop = np.zeros((N * T, ldh), 'float32')
    dop[:, :H] = dout.reshape(N * T, H)
    d_dout = G.dev(dop)
    d_dg, d_dpeep = G.zeros((N * T, 4 * H)), (G.zeros((3, H)) if peep else None)
    d_dci, d_dhi = G.zeros((H,)), G.zeros((H,))
    G.call('ipavsr_lstm_bwd', d_dout.data_ptr(), d_whid.data_ptr(), G.ptr(d_peep), d_ci.data_ptr(), d_mask.data_ptr(),
           d_gates.data_ptr(), d_cell.data_ptr(), d_dg.data_ptr(), G.ptr(d_dpeep), d_dci.data_ptr(), d_dhi.data_ptr(),
           N, T, H, ldh, int(backwards), 5.0, 0, impl, ws.data_ptr(), nbytes, G.stream())
    dG = G.deinterleave_gates(G.host(d_dg), H).astype(np.float64)          # (N*T, 4H) in [i|f|c|o] order
    # derived gradients exactly as the engine derives them
    dW_in = x.reshape(N * T, I).astype(np.float64).T @ dG
    hprev = G.host(d_hprev)[:, :H].astype(np.float64)
    dW_hid = hprev.T @ dG
    db = dG.sum(0)
    dx = (dG @ p['W_in'].astype(np.float64).T).reshape(N, T, I)
    tol = 3e-4
    assert G.relerr(dW_in, gr['W_in']) < tol, ('W_in', G.relerr(dW_in, gr['W_in']))
    assert G.relerr(dW_hid, gr['W_hid']) < tol, ('W_hid', G.relerr(dW_hid, gr['W_hid']))
    assert G.relerr(db, gr['b']) < tol
    assert G.relerr(dx, dx_ref) < tol
    assert G.relerr(G.host(d_dci), gr['cell_init']) < tol, ('cell_init', G.relerr(G.host(d_dci), gr['cell_init']))
    assert G.relerr(G.host(d_dhi), gr['hid_init']) < tol, ('hid_init', G.relerr(G.host(d_dhi), gr['hid_init']))
    if peep:
        assert G.relerr(G.host(d_dpeep), gr['peep']) < tol, ('peep', G.relerr(G.host(d_dpeep), gr['peep']))


@pytest.mark.parametrize('N,T,H,peep,backwards', [(26, 40, 250, True, False), (70, 40, 250, False, True), (5, 7, 40, True, True),
                                                    (33, 3, 8, True, False), (512, 40, 250, True, False), (40, 64, 256, True, True),
                                                    (9, 1, 100, False, False)])
def test_lstm_fwd_tensor_core(N, T, H, peep, backwards):
    """ipavsr_lstm_fwd_f16 (tcgen05 recurrence on the fp16 split of W_hid) == oracle, including the training saves."""
    rng = np.random.default_rng(N + T + H)
    lens = rng.integers(1, T + 1, size=N)
    lens[0] = T
    p, x, mask = _lstm_inputs(rng, N, T, 12, H, peep, lens)
    p['hid_init'] = (p['hid_init'] * (4.0 if backwards else 1.0)).astype('float32')       # |hid_init| > 1 too
    if H >= 100:
        # N(0, 0.3) at H = 250 is an expanding recurrence (gain ~1.2 per step) that amplifies the ~2e-7 per-step error of
        # the three-product tensor-core arithmetic (truncating TMEM accumulation) to ~1e-4 over 40 steps; the models use
        # orthogonal recurrent weights (gain <= 0.25 per step).  Keep the recurrence non-expanding here.
        p['W_hid'] = (p['W_hid'] * 0.25).astype('float32')
    out_ref, cache = ops.lstm_fwd(x, mask, p, backwards, np.float64)
    xw = (x.reshape(N * T, 12).astype(np.float64) @ p['W_in'].astype(np.float64) + p['b']).astype('float32')
    ldh = (H + 7) // 8 * 8
    ldw = (4 * H + 7) // 8 * 8
    whid = np.zeros((H, ldw), 'float32')
    whid[:, :4 * H] = G.interleave_gates(p['W_hid'], H)
    d_xw, d_whid = G.dev(G.interleave_gates(xw, H)), G.dev(whid)
    wh, wl = G.zeros((H, ldw), torch.float16), G.zeros((H, ldw), torch.float16)
    sc = G.zeros((2,))
    G.call('ipavsr_f16_split', d_whid.data_ptr(), ldw, H, 4 * H, wh.data_ptr(), wl.data_ptr(), ldw, sc.data_ptr(),
           sc.data_ptr() + 4, 0, G.stream())
    d_peep = G.dev(p['peep']) if peep else None
    d_ci, d_hi, d_mask = G.dev(p['cell_init']), G.dev(p['hid_init']), G.dev(mask)
    d_out, d_gates = G.zeros((N * T, ldh)), G.zeros((N * T, 4 * H))
    d_cell, d_hprev = G.zeros((N * T, H)), G.zeros((N * T, ldh))
    assert G.lib().ipavsr_lstm_fwd_f16_supported(N, T, H, ldw)
    G.call('ipavsr_lstm_fwd_f16', d_xw.data_ptr(), wh.data_ptr(), wl.data_ptr(), sc.data_ptr() + 4, ldw, G.ptr(d_peep),
           d_ci.data_ptr(), d_hi.data_ptr(), d_mask.data_ptr(), d_out.data_ptr(), d_gates.data_ptr(), d_cell.data_ptr(),
           d_hprev.data_ptr(), N, T, H, ldh, int(backwards), G.stream())
    out = G.host(d_out)[:, :H].reshape(N, T, H)
    tol = 2e-5
    assert G.relerr(out, out_ref) < tol, G.relerr(out, out_ref)
    # training saves agree with the FFMA kernel's
    e_out, e_gates = G.zeros((N * T, ldh)), G.zeros((N * T, 4 * H))
    e_cell, e_hprev = G.zeros((N * T, H)), G.zeros((N * T, ldh))
    nbytes = G.lib().ipavsr_lstm_workspace_bytes(N, T, H)
    ws = G.zeros(((nbytes + 3) // 4,))
    d_w32 = G.dev(G.interleave_gates(p['W_hid'], H))
    G.call('ipavsr_lstm_fwd', d_xw.data_ptr(), d_w32.data_ptr(), G.ptr(d_peep), d_ci.data_ptr(), d_hi.data_ptr(),
           d_mask.data_ptr(), e_out.data_ptr(), e_gates.data_ptr(), e_cell.data_ptr(), e_hprev.data_ptr(), N, T, H, ldh,
           int(backwards), 0, ws.data_ptr(), nbytes, G.stream())
    # the gate activations of a masked step are never read (the state passes through); the tensor-core kernel does not
    # compute them at frames where its whole 32-utterance tile is masked, so they are compared at unmasked steps only
    live = mask.reshape(-1).astype(bool)
    assert G.relerr(G.host(d_gates)[live], G.host(e_gates)[live]) < tol, 'gates'
    for a, b, name in ((d_cell, e_cell, 'cell'), (d_hprev, e_hprev, 'hprev'), (d_out, e_out, 'out')):
        assert G.relerr(G.host(a), G.host(b)) < tol, name


@pytest.mark.parametrize('N,T,H,peep,backwards,scale', [(26, 40, 250, True, False, 1.0), (70, 40, 250, False, True, 30.0),
                                                          (5, 7, 40, True, True, 1.0), (33, 3, 8, True, False, 1.0),
                                                          (200, 12, 256, True, True, 1.0), (9, 1, 100, False, False, 1.0)])
def test_lstm_bwd_tensor_core(N, T, H, peep, backwards, scale):
    """ipavsr_lstm_bwd_f16 (tcgen05 BPTT) == the FFMA kernel and the oracle on the same saved tensors."""
    rng = np.random.default_rng(N + 2 * T + H)
    lens = rng.integers(1, T + 1, size=N)
    lens[0] = T
    I = 12
    p, x, mask = _lstm_inputs(rng, N, T, I, H, peep, lens)
    if H >= 100:
        p['W_hid'] = (p['W_hid'] * 0.25).astype('float32')
    dout = (rng.normal(size=(N, T, H)) * scale).astype('float32')
    out_ref, cache = ops.lstm_fwd(x, mask, p, backwards, np.float64)
    dx_ref, gr = ops.lstm_bwd(dout, cache, 5.0, np.float64)
    xw = (x.reshape(N * T, I).astype(np.float64) @ p['W_in'].astype(np.float64) + p['b']).astype('float32')
    ldh, ldw = (H + 7) // 8 * 8, 4 * H
    d_xw = G.dev(G.interleave_gates(xw, H))
    d_whid = G.dev(G.interleave_gates(p['W_hid'], H))
    wh, wl, sc = G.zeros((H, ldw), torch.float16), G.zeros((H, ldw), torch.float16), G.zeros((2,))
    G.call('ipavsr_f16_split', d_whid.data_ptr(), ldw, H, 4 * H, wh.data_ptr(), wl.data_ptr(), ldw, sc.data_ptr(),
           sc.data_ptr() + 4, 0, G.stream())
    d_peep = G.dev(p['peep']) if peep else None
    d_ci, d_hi, d_mask = G.dev(p['cell_init']), G.dev(p['hid_init']), G.dev(mask)
    d_out, d_gates = G.zeros((N * T, ldh)), G.zeros((N * T, 4 * H))
    d_cell, d_hprev = G.zeros((N * T, H)), G.zeros((N * T, ldh))
    nbytes = G.lib().ipavsr_lstm_workspace_bytes(N, T, H)
    ws = G.zeros(((nbytes + 3) // 4,))
    G.call('ipavsr_lstm_fwd', d_xw.data_ptr(), d_whid.data_ptr(), G.ptr(d_peep), d_ci.data_ptr(), d_hi.data_ptr(),
           d_mask.data_ptr(), d_out.data_ptr(), d_gates.data_ptr(), d_cell.data_ptr(), d_hprev.data_ptr(), N, T, H, ldh,
           int(backwards), 0, ws.data_ptr(), nbytes, G.stream())
    dop = np.zeros((N * T, ldh), 'float32')
    dop[:, :H] = dout.reshape(N * T, H)
    d_dout = G.dev(dop)
    res = []
    d_db = G.zeros((4 * H,))
    dgh, dgl = G.zeros((N * T, 4 * H), torch.float16), G.zeros((N * T, 4 * H), torch.float16)
    dge = G.zeros((1,), torch.int32)
    for tc in (True, False):
        d_dg, d_dpeep = G.zeros((N * T, 4 * H)), (G.zeros((3, H)) if peep else None)
        d_dci, d_dhi = G.zeros((H,)), G.zeros((H,))
        if tc:
            assert G.lib().ipavsr_lstm_bwd_f16_supported(N, T, H, ldw, 5.0)
            G.call('ipavsr_lstm_bwd_f16', d_dout.data_ptr(), d_whid.data_ptr(), wh.data_ptr(), wl.data_ptr(), sc.data_ptr() + 4,
                   ldw, G.ptr(d_peep), d_ci.data_ptr(), d_mask.data_ptr(), d_gates.data_ptr(), d_cell.data_ptr(),
                   d_dg.data_ptr(), G.ptr(d_dpeep), d_dci.data_ptr(), d_dhi.data_ptr(), N, T, H, ldh, int(backwards), 5.0, 0,
                   d_db.data_ptr(), dgh.data_ptr(), dgl.data_ptr(), dge.data_ptr(), ws.data_ptr(), nbytes, G.stream())
            # by-products: bias gradient and the fp16 split of dgates
            dgv = G.host(d_dg).astype(np.float64)
            assert G.relerr(G.host(d_db), dgv.sum(0)) < 1e-5
            ex = int(G.host(dge)[0])
            rec = (G.host(dgh).astype(np.float64) + G.host(dgl).astype(np.float64) / G.F16_LO_SCALE) / 2.0 ** ex
            assert np.abs(rec - dgv).max() <= max(np.abs(dgv).max() * 2.0 ** -20, 1e-12)
        else:
            G.call('ipavsr_lstm_bwd', d_dout.data_ptr(), d_whid.data_ptr(), G.ptr(d_peep), d_ci.data_ptr(), d_mask.data_ptr(),
                   d_gates.data_ptr(), d_cell.data_ptr(), d_dg.data_ptr(), G.ptr(d_dpeep), d_dci.data_ptr(), d_dhi.data_ptr(),
                   N, T, H, ldh, int(backwards), 5.0, 0, 0, ws.data_ptr(), nbytes, G.stream())
        res.append((G.host(d_dg), G.host(d_dci), G.host(d_dhi), G.host(d_dpeep) if peep else None))
    tol = 3e-4
    for a, b, name in zip(res[0], res[1], ('dgates', 'dcell_init', 'dhid_init', 'dpeep')):
        if a is not None:
            assert G.relerr(a, b) < tol, (name, G.relerr(a, b))
    dG = G.deinterleave_gates(res[0][0], H).astype(np.float64)
    assert G.relerr(dG.sum(0), gr['b']) < tol
    assert G.relerr((dG @ p['W_in'].astype(np.float64).T).reshape(N, T, I), dx_ref) < tol
    assert G.relerr(res[0][1], gr['cell_init']) < tol and G.relerr(res[0][2], gr['hid_init']) < tol


def test_lstm_fwd_tensor_core_orthogonal_weights():
    """With the models' own initialisation (orthogonal W_hid, lasagne.init.Orthogonal) the tensor-core recurrence holds the
    1e-5 gate of the FFMA kernel over T = 40."""
    rng = np.random.default_rng(5)
    N, T, H = 64, 40, 250
    lens = rng.integers(12, T + 1, size=N)
    p, x, mask = _lstm_inputs(rng, N, T, 150, H, True, lens)
    for g in range(4):
        q, _ = np.linalg.qr(rng.normal(size=(H, H)))
        p['W_hid'][:, g * H:(g + 1) * H] = q.astype('float32')
    p['W_in'] = (p['W_in'] / np.sqrt(150) / 0.3).astype('float32')
    out_ref, _ = ops.lstm_fwd(x, mask, p, False, np.float64)
    xw = (x.reshape(N * T, 150).astype(np.float64) @ p['W_in'].astype(np.float64) + p['b']).astype('float32')
    ldh, ldw = 256, 1000
    d_xw, d_whid = G.dev(G.interleave_gates(xw, H)), G.dev(G.interleave_gates(p['W_hid'], H))
    wh, wl, sc = G.zeros((H, ldw), torch.float16), G.zeros((H, ldw), torch.float16), G.zeros((2,))
    G.call('ipavsr_f16_split', d_whid.data_ptr(), ldw, H, 4 * H, wh.data_ptr(), wl.data_ptr(), ldw, sc.data_ptr(),
           sc.data_ptr() + 4, 0, G.stream())
    d_out = G.zeros((N * T, ldh))
    d_peep, d_ci, d_hi, d_mask = G.dev(p['peep']), G.dev(p['cell_init']), G.dev(p['hid_init']), G.dev(mask)
    G.call('ipavsr_lstm_fwd_f16', d_xw.data_ptr(), wh.data_ptr(), wl.data_ptr(), sc.data_ptr() + 4, ldw,
           d_peep.data_ptr(), d_ci.data_ptr(), d_hi.data_ptr(), d_mask.data_ptr(), d_out.data_ptr(), None, None, None,
           N, T, H, ldh, 0, G.stream())
    err = G.relerr(G.host(d_out)[:, :H].reshape(N, T, H), out_ref)
    assert err < 1e-5, err


def test_lstm_impls_agree_large():
    """persistent cluster kernel == step-wise form at a batch that spans several cluster tiles."""
    rng = np.random.default_rng(77)
    N, T, H = 200, 40, 250
    lens = rng.integers(10, T + 1, size=N)
    xw = rng.normal(0, 1, (N * T, 4 * H)).astype('float32')
    whid = rng.normal(0, .1, (H, 4 * H)).astype('float32')
    peep = rng.normal(0, .1, (3, H)).astype('float32')
    mask = (np.arange(T)[None, :] < lens[:, None]).astype('uint8')
    ldh = 252
    outs = []
    d_xw, d_whid, d_peep, d_mask = G.dev(xw), G.dev(whid), G.dev(peep), G.dev(mask)
    for impl in (0, 1):
        d_out = G.zeros((N * T, ldh))
        nbytes = G.lib().ipavsr_lstm_workspace_bytes(N, T, H)
        ws = G.zeros(((nbytes + 3) // 4,))
        z = G.zeros((H,))
        G.call('ipavsr_lstm_fwd', d_xw.data_ptr(), d_whid.data_ptr(), d_peep.data_ptr(), z.data_ptr(),
               z.data_ptr(), d_mask.data_ptr(), d_out.data_ptr(), None, None, None, N, T, H, ldh, 0, impl,
               ws.data_ptr(), nbytes, G.stream())
        outs.append(G.host(d_out))
    assert G.relerr(outs[0], outs[1]) < 1e-5


# ---------------------------------------------------------------------------------------------------------
# elementwise / reductions / BN / losses / optimiser
# ---------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize('act', [1, 2, 3, 4, 6, 7])
def test_dense_bwd_prep_and_colsum(act):
    rng = np.random.default_rng(act)
    M, N = 1037, 203
    z = rng.normal(size=(M, N))
    y = ops.act_fwd(z, act).astype('float32')
    dy = rng.normal(size=(M, N)).astype('float32')
    want = ops.act_bwd(dy.astype(np.float64), z, y.astype(np.float64), act)
    ld = 204
    pad = lambda a: np.pad(a, ((0, 0), (0, ld - N)))
    d_dy, d_y, d_dz, d_db = G.dev(pad(dy)), G.dev(pad(y)), G.zeros((M, ld)), G.zeros((N,))
    d_amax = G.zeros((1,))
    G.call('ipavsr_dense_bwd_prep', d_dy.data_ptr(), ld, d_y.data_ptr(), ld, d_dz.data_ptr(), ld, d_db.data_ptr(), M, N,
           act, 0, d_amax.data_ptr(), G.stream())
    assert G.relerr(G.host(d_dz)[:, :N], want) < 3e-6
    assert G.host(d_amax)[0] == np.abs(G.host(d_dz)[:, :N]).max()
    assert G.relerr(G.host(d_db), want.sum(0)) < 1e-5
    G.call('ipavsr_colsum', d_dy.data_ptr(), ld, d_db.data_ptr(), M, N, 0, G.stream())
    assert G.relerr(G.host(d_db), dy.astype(np.float64).sum(0)) < 1e-5


def test_fuse_sum_adasum_copy_slice_dropout():
    rng = np.random.default_rng(3)
    M, F, S = 520, 250, 3
    ld = 252
    xs = [rng.normal(size=(M, ld)).astype('float32') for _ in range(S)]
    co = rng.normal(size=(S,)).astype('float32')
    import ctypes as C
    dxs = [G.dev(x) for x in xs]
    ptrs = (C.c_void_p * S)(*[d.data_ptr() for d in dxs])
    lds = (C.c_int * S)(*[ld] * S)
    out = G.zeros((M, ld))
    G.call('ipavsr_fuse_sum', ptrs, lds, S, None, out.data_ptr(), ld, M, F, G.stream())
    assert G.relerr(G.host(out)[:, :F], sum(x[:, :F] for x in xs)) < 1e-6
    dco = G.dev(co)
    G.call('ipavsr_fuse_sum', ptrs, lds, S, dco.data_ptr(), out.data_ptr(), ld, M, F, G.stream())
    assert G.relerr(G.host(out)[:, :F], sum(c * x[:, :F] for c, x in zip(co, xs))) < 1e-6
    g = rng.normal(size=(M, ld)).astype('float32')
    dcoef = G.zeros((S,))
    d_g = G.dev(g)
    G.call('ipavsr_adasum_bwd_coeff', d_g.data_ptr(), ld, ptrs, lds, S, dcoef.data_ptr(), M, F, 0, G.stream())
    want = [(g[:, :F].astype(np.float64) * x[:, :F]).sum() for x in xs]
    assert G.relerr(G.host(dcoef), want) < 1e-4
    # copy2d with device alpha + accumulate
    dst = G.dev(xs[1])
    G.call('ipavsr_copy2d', dxs[0].data_ptr(), ld, dst.data_ptr(), ld, M, F, dco.data_ptr() + 4, 1, G.stream())
    assert G.relerr(G.host(dst)[:, :F], xs[1][:, :F] + co[1] * xs[0][:, :F]) < 1e-6
    # slice last (N=13, T=40)
    N, T = 13, 40
    sl = G.zeros((N, ld))
    G.call('ipavsr_slice_last', dxs[2].data_ptr(), ld, sl.data_ptr(), ld, N, T, F, 0, 0, G.stream())
    np.testing.assert_array_equal(G.host(sl)[:, :F], xs[2].reshape(N, T, ld)[:, -1, :F])
    back = G.zeros((M, ld))
    G.call('ipavsr_slice_last', sl.data_ptr(), ld, back.data_ptr(), ld, N, T, F, 1, 0, G.stream())
    b = G.host(back).reshape(N, T, ld)
    np.testing.assert_array_equal(b[:, -1, :F], xs[2].reshape(N, T, ld)[:, -1, :F])
    assert np.abs(b[:, :-1]).max() == 0
    # dropout with explicit mask, and the generator's keep rate
    keep = (rng.random((M, F)) > 0.5).astype('uint8')
    y = G.zeros((M, ld))
    d_keep = G.dev(keep)
    G.call('ipavsr_dropout', dxs[0].data_ptr(), ld, d_keep.data_ptr(), y.data_ptr(), ld, M, F, 2.0, G.stream())
    np.testing.assert_allclose(G.host(y)[:, :F], xs[0][:, :F] * keep * 2.0, rtol=1e-7)
    km = G.zeros((1 << 20,), torch.uint8)
    G.call('ipavsr_dropout_mask', km.data_ptr(), 1 << 20, 0.2, 42, 0, G.stream())
    assert abs(G.host(km).mean() - 0.8) < 5e-3


def test_batchnorm():
    rng = np.random.default_rng(8)
    M, F = 1040, 50
    ld = 52
    x = rng.normal(1.0, 2.0, size=(M, F)).astype('float32')
    beta, gamma = rng.normal(size=F).astype('float32'), rng.normal(1, .2, size=F).astype('float32')
    dy = rng.normal(size=(M, F)).astype('float32')
    y_ref, cache, new = ops.bn_fwd(x, beta, gamma, np.zeros(F), np.ones(F), False, 1e-4, 0.1, np.float64)
    dx_ref, dbeta_ref, dgamma_ref = ops.bn_bwd(dy, cache, gamma, np.float64)
    pad = lambda a: np.pad(a, ((0, 0), (0, ld - F)))
    d_x, d_y = G.dev(pad(x)), G.zeros((M, ld))
    stats, save = G.zeros((2 * F,), torch.float64), G.zeros((2 * F,))
    rm, ri = G.zeros((F,)), G.dev(np.ones(F, 'float32'))
    d_beta, d_gamma = G.dev(beta), G.dev(gamma)
    G.call('ipavsr_bn_stats', d_x.data_ptr(), ld, stats.data_ptr(), M, F, G.stream())
    G.call('ipavsr_bn_fwd', d_x.data_ptr(), ld, d_y.data_ptr(), ld, d_beta.data_ptr(), d_gamma.data_ptr(), rm.data_ptr(),
           ri.data_ptr(), stats.data_ptr(), save.data_ptr(), save.data_ptr() + 4 * F, M, F, M, 1e-4, 0.1, 0, 1,
           G.stream())
    assert G.relerr(G.host(d_y)[:, :F], y_ref) < 3e-6
    assert G.relerr(G.host(rm), new[0]) < 1e-5 and G.relerr(G.host(ri), new[1]) < 1e-5
    bst = G.zeros((2 * F,), torch.float64)
    d_dy, d_dx, d_db, d_dg = G.dev(pad(dy)), G.zeros((M, ld)), G.zeros((F,)), G.zeros((F,))
    G.call('ipavsr_bn_bwd_stats', d_dy.data_ptr(), ld, d_x.data_ptr(), ld, save.data_ptr(), save.data_ptr() + 4 * F,
           bst.data_ptr(), M, F, G.stream())
    G.call('ipavsr_bn_bwd', d_dy.data_ptr(), ld, d_x.data_ptr(), ld, d_gamma.data_ptr(), save.data_ptr(),
           save.data_ptr() + 4 * F, bst.data_ptr(), d_dx.data_ptr(), ld, d_db.data_ptr(), d_dg.data_ptr(), M, F, M, 0,
           G.stream())
    assert G.relerr(G.host(d_dx)[:, :F], dx_ref) < 1e-5
    assert G.relerr(G.host(d_db), dbeta_ref) < 1e-5 and G.relerr(G.host(d_dg), dgamma_ref) < 1e-5
    # deterministic mode uses the running statistics
    y2, _, _ = ops.bn_fwd(x, beta, gamma, G.host(rm), G.host(ri), True, 1e-4, 0.1, np.float64)
    G.call('ipavsr_bn_fwd', d_x.data_ptr(), ld, d_y.data_ptr(), ld, d_beta.data_ptr(), d_gamma.data_ptr(), rm.data_ptr(),
           ri.data_ptr(), None, None, None, M, F, 0, 1e-4, 0.1, 1, 0, G.stream())
    assert G.relerr(G.host(d_y)[:, :F], y2) < 3e-6


def test_softmax_and_losses():
    rng = np.random.default_rng(12)
    N, T, Cc = 26, 40, 26
    ld = 28
    z = rng.normal(0, 2, size=(N * T, Cc)).astype('float32')
    y = rng.integers(0, Cc, size=(N, T)).astype('int32')
    lens = rng.integers(1, T + 1, size=N)
    mask = (np.arange(T)[None] < lens[:, None]).astype('uint8')
    pad = lambda a: np.pad(a, ((0, 0), (0, ld - Cc)))
    d_z, d_p = G.dev(pad(z)), G.zeros((N * T, ld))
    G.call('ipavsr_softmax', d_z.data_ptr(), ld, d_p.data_ptr(), ld, N * T, Cc, G.stream())
    p_ref = ops.softmax_rows(z.astype(np.float64))
    assert G.relerr(G.host(d_p)[:, :Cc], p_ref) < 2e-6
    loss_ref, dp_ref = ops.temporal_softmax_loss(p_ref.reshape(N, T, Cc), y, mask, np.float64)
    dz_ref = ops.act_bwd(dp_ref.reshape(N * T, Cc), None, p_ref, ops.ACT_SOFTMAX)
    ls, dl = G.zeros((4,)), G.zeros((N * T, ld))
    d_yy, d_mm = G.dev(y), G.dev(mask)
    cnt = float(mask.sum())
    G.call('ipavsr_temporal_softmax_loss', d_p.data_ptr(), ld, d_yy.data_ptr(), d_mm.data_ptr(),
           ls.data_ptr(), dl.data_ptr(), ld, N * T, Cc, 1.0 / cnt, None, G.stream())
    assert abs(G.host(ls)[0] / cnt - loss_ref) < 1e-5 * abs(loss_ref)
    assert G.relerr(G.host(dl)[:, :Cc], dz_ref) < 1e-5
    # device-side normaliser gives the same gradient
    cd = G.dev(np.array([cnt], 'float32'))
    dl2 = G.zeros((N * T, ld))
    ls.zero_()
    G.call('ipavsr_temporal_softmax_loss', d_p.data_ptr(), ld, d_yy.data_ptr(), d_mm.data_ptr(),
           ls.data_ptr(), dl2.data_ptr(), ld, N * T, Cc, 1.0, cd.data_ptr(), G.stream())
    assert G.relerr(G.host(dl2), G.host(dl)) < 1e-6
    # sequence-level
    yv = rng.integers(0, Cc, size=(N,)).astype('int32')
    pn = p_ref[:N]
    loss2, dp2 = ops.categorical_crossentropy_mean(pn, yv, np.float64)
    dz2 = ops.act_bwd(dp2, None, pn, ops.ACT_SOFTMAX)
    ls.zero_()
    dl3 = G.zeros((N, ld))
    d_yv = G.dev(yv)
    G.call('ipavsr_categorical_crossentropy', d_p.data_ptr(), ld, d_yv.data_ptr(), ls.data_ptr(), dl3.data_ptr(),
           ld, N, Cc, 1.0 / N, None, G.stream())
    assert abs(G.host(ls)[0] / N - loss2) < 1e-5 * abs(loss2)
    assert G.relerr(G.host(dl3)[:, :Cc], dz2) < 1e-5


@pytest.mark.parametrize('kind', ['adam', 'adadelta', 'sgd', 'momentum', 'nesterov'])
def test_optim_step(kind):
    rng = np.random.default_rng(21)
    n = 256 * 37
    p0 = rng.normal(size=n).astype('float32')
    gs = [rng.normal(size=n).astype('float32') for _ in range(4)]
    p = p0.copy()
    d_p, s1, s2 = G.dev(p0), G.zeros((n,)), G.zeros((n,))
    code = {'adam': 0, 'adadelta': 1, 'sgd': 2, 'momentum': 3, 'nesterov': 4}[kind]
    st = {'t': np.float32(0), 'm': [np.zeros(n, 'f')], 'v': [np.zeros(n, 'f')], 'acc': [np.zeros(n, 'f')],
          'dacc': [np.zeros(n, 'f')], 'vel': [np.zeros(n, 'f')]}
    t = np.float32(0)
    for g in gs:
        if kind == 'adam':
            ops.adam_step([p], [g], st, [1e-3])
            t = np.float32(t + 1)
            sc = float(np.sqrt(np.float32(1) - np.float32(0.999) ** t) / (np.float32(1) - np.float32(0.9) ** t))
            args = (1e-3, None, None, sc, 0.9, 0.999, 1e-8, 1.0)
        elif kind == 'adadelta':
            ops.adadelta_step([p], [g], st, 2.0)
            args = (2.0, None, None, 1.0, 0.95, 0.0, 1e-6, 1.0)
        elif kind == 'sgd':
            ops.sgd_step([p], [g], 0.01)
            args = (0.01, None, None, 1.0, 0.0, 0.0, 0.0, 1.0)
        else:
            ops.sgd_momentum_step([p], [g], st, 0.01, 0.9, kind == 'nesterov')
            args = (0.01, None, None, 1.0, 0.9, 0.0, 0.0, 1.0)
        d_g = G.dev(g)
        G.call('ipavsr_optim_step', code, d_p.data_ptr(), d_g.data_ptr(), s1.data_ptr(), s2.data_ptr(), n, *args,
               G.stream())
    np.testing.assert_allclose(G.host(d_p), p, rtol=2e-5, atol=2e-6)


def test_optim_step_variable_lr():
    rng = np.random.default_rng(22)
    n = 256 * 8
    p0 = rng.normal(size=n).astype('float32')
    g = rng.normal(size=n).astype('float32')
    seg_id = np.repeat(np.arange(8) % 3, 1).astype('int32')
    seg_lr = np.array([1e-3, 0.0, 5e-3], 'float32')
    d_p = G.dev(p0)
    d_g, d_sl, d_si = G.dev(g), G.dev(seg_lr), G.dev(seg_id)
    G.call('ipavsr_optim_step', 2, d_p.data_ptr(), d_g.data_ptr(), None, None, n, 0.0, d_sl.data_ptr(),
           d_si.data_ptr(), 1.0, 0.0, 0.0, 0.0, 1.0, G.stream())
    lr_full = np.repeat(seg_lr[seg_id], 256)
    np.testing.assert_allclose(G.host(d_p), p0 - lr_full * g, rtol=1e-6, atol=1e-7)


# ---------------------------------------------------------------------------------------------------------
# utils/preprocessing on device vs golden vectors produced by the reference itself
# ---------------------------------------------------------------------------------------------------------
def test_preprocessing_golden():
    X, lens = GOLD['X'], GOLD['lens']
    frames, D = X.shape
    offs = G.dev(np.concatenate([[0], np.cumsum(lens)]).astype('int64'))
    d_x, d_y = G.dev(X), G.zeros((frames, D))
    G.call('ipavsr_norm_samplewise', d_x.data_ptr(), D, d_y.data_ptr(), D, frames, D, G.stream())
    np.testing.assert_allclose(G.host(d_y), GOLD['normalize_input'], rtol=2e-6, atol=2e-6)
    mean, std, scratch = G.zeros((D,)), G.zeros((D,)), G.zeros((3 * D,), torch.float64)
    G.call('ipavsr_norm_featurewise_stats', d_x.data_ptr(), D, mean.data_ptr(), std.data_ptr(), scratch.data_ptr(),
           frames, D, G.stream())
    np.testing.assert_allclose(G.host(mean), GOLD['featurewise_mean'], rtol=1e-6, atol=1e-6)
    np.testing.assert_allclose(G.host(std), GOLD['featurewise_std'], rtol=1e-6)
    G.call('ipavsr_norm_featurewise_apply', d_x.data_ptr(), D, mean.data_ptr(), std.data_ptr(), d_y.data_ptr(), D,
           frames, D, G.stream())
    np.testing.assert_allclose(G.host(d_y), GOLD['featurewise_norm'], rtol=1e-5, atol=2e-6)
    G.call('ipavsr_seq_mean_sub', d_x.data_ptr(), D, d_y.data_ptr(), D, offs.data_ptr(), len(lens), D, G.stream())
    np.testing.assert_array_equal(G.host(d_y), GOLD['seq_mean_sub'])            # same summation order: bit-exact
    G.call('ipavsr_diff_image', d_x.data_ptr(), D, d_y.data_ptr(), D, offs.data_ptr(), len(lens), D, G.stream())
    np.testing.assert_array_equal(G.host(d_y), GOLD['diff_images'])
    Xd = GOLD['Xd']
    F = Xd.shape[1]
    for w in (9, 5):
        out = G.zeros((frames, 3 * F), torch.float64)
        d_xd = G.dev(Xd)
        G.call('ipavsr_deltas_fir', d_xd.data_ptr(), F, out.data_ptr(), 3 * F, offs.data_ptr(), len(lens), F, w,
               int(lens.max()), G.stream())
        np.testing.assert_allclose(G.host(out), GOLD['concat_deltas_w%d' % w], rtol=1e-12, atol=1e-10)


def test_preprocessing_large_properties():
    """Full-size style checks through size-independent properties (SURVEY §8d config 4 shapes, reduced frames)."""
    rng = np.random.default_rng(31)
    U, T, D = 512, 40, 1200
    frames = U * T
    X = rng.normal(2, 3, size=(frames, D)).astype('float32')
    d_x, d_y = G.dev(X), G.zeros((frames, D))
    G.call('ipavsr_norm_samplewise', d_x.data_ptr(), D, d_y.data_ptr(), D, frames, D, G.stream())
    y = G.host(d_y)
    assert np.abs(y.mean(1)).max() < 1e-5 and np.abs(y.std(1) - 1).max() < 1e-5
    G.call('ipavsr_norm_samplewise', d_y.data_ptr(), D, d_x.data_ptr(), D, frames, D, G.stream())   # idempotent
    assert np.abs(G.host(d_x) - y).max() < 1e-5
    offs = G.dev((np.arange(U + 1) * T).astype('int64'))
    G.call('ipavsr_seq_mean_sub', d_y.data_ptr(), D, d_x.data_ptr(), D, offs.data_ptr(), U, D, G.stream())
    z = G.host(d_x).reshape(U, T, D)
    assert np.abs(z.mean(1)).max() < 1e-5
    G.call('ipavsr_diff_image', d_y.data_ptr(), D, d_x.data_ptr(), D, offs.data_ptr(), U, D, G.stream())
    dd = G.host(d_x).reshape(U, T, D)
    yy = y.reshape(U, T, D)
    np.testing.assert_array_equal(dd[:, 1:], yy[:, 1:] - yy[:, :-1])
    np.testing.assert_array_equal(dd[:, 0], dd[:, 1])
    # oracle on a slice
    np.testing.assert_allclose(y[:64], OP.normalize_input(X[:64]), rtol=3e-6, atol=3e-6)


def test_preprocessing_python_api_matches_reference_golden():
    """The reference-named Python functions (ipavsr_b200.utils.preprocessing) against the reference's own outputs."""
    from ipavsr_b200.utils import preprocessing as P, signal as S
    X, lens = GOLD['X'], GOLD['lens']
    np.testing.assert_allclose(P.normalize_input(X.copy()), GOLD['normalize_input'], rtol=2e-6, atol=2e-6)
    n, m, s = P.featurewise_normalize_sequence(X)
    np.testing.assert_allclose(n, GOLD['featurewise_norm'], rtol=1e-5, atol=2e-6)
    np.testing.assert_allclose(m, GOLD['featurewise_mean'], rtol=1e-6, atol=1e-6)
    np.testing.assert_array_equal(P.sequencewise_mean_image_subtraction(X, lens), GOLD['seq_mean_sub'])
    np.testing.assert_array_equal(P.compute_diff_images(X, lens), GOLD['diff_images'])
    np.testing.assert_allclose(P.concat_first_second_deltas(GOLD['Xd'], lens, 9), GOLD['concat_deltas_w9'], rtol=1e-12,
                               atol=1e-10)
    np.testing.assert_allclose(P.deltas(GOLD['test_delta_in'], 9), GOLD['test_delta_out'], atol=1e-9)
    np.testing.assert_allclose(P.deltas(GOLD['leftpad_in'], 9), GOLD['leftpad_out'], atol=1e-9)
    with pytest.raises(IndexError):
        P.compute_diff_images(X[:3], [1, 2])
    seqs = np.array([[[1, 2, 3, 4, 5], [10, 12, 13, 14, 15], [300, 1, 23, 56, 22]]], dtype='float32')
    np.testing.assert_array_equal(S.append_delta_coeff(seqs[0], 1)[0],
                                  [1, 2, 3, 4, 5, 4.5, 5, 5, 5, 5, 72.5, -2.75, 2.5, 10.5, 1.75])


def test_delta_bwd_alternating_windows():
    """The register-column backward keeps its boundary weights in a constant table keyed by Theta: alternating window
    sizes (and frame counts) must re-upload it every time it changes."""
    rng = np.random.default_rng(3)
    for theta, T in ((9, 20), (4, 20), (9, 20), (1, 33), (9, 40), (4, 40), (9, 40)):
        N, F = 7, 10
        gy = rng.normal(size=(N, T, 3 * F)).astype('float32')
        want = ops.delta_bwd(gy, theta, np.float64)
        d_gy, d_gx = G.dev(np.pad(gy.reshape(N * T, 3 * F), ((0, 0), (0, 2)))), G.zeros((N * T, 16))
        G.call('ipavsr_delta_bwd', d_gy.data_ptr(), 3 * F + 2, d_gx.data_ptr(), 16, N, T, F, theta, 0, G.stream())
        got = G.host(d_gx)[:, :F].reshape(N, T, F)
        assert G.relerr(got, want) < 1e-5, (theta, T)


def test_lstm_drift_expanding_recurrence():
    """Trained recurrent weights are not orthogonal.  With W_hid ~ N(0, 0.3) at H = 250 the recurrence EXPANDS perturbations
    (the Jacobian's gain per step is > 1), so any float32-class arithmetic drifts away from the float64 result over T = 40
    steps — the CUDA-core FFMA kernel, NumPy float32 and the three-product tensor-core kernel alike.  The stated bound for
    this regime: the tensor-core recurrence stays within 4x the drift of the FFMA float32 kernel on the same weights (and
    below 1e-3 relative), i.e. it is in the same accuracy class; with non-expanding weights (the other tests) both hold 2e-5."""
    N, T, H, I = 64, 40, 250, 12
    rng = np.random.default_rng(11)
    p, x, mask = _lstm_inputs(rng, N, T, I, H, True, np.full(N, T))
    ref64, _ = ops.lstm_fwd(x, mask, p, False, np.float64)
    ref32, _ = ops.lstm_fwd(x, mask, p, False, np.float32)
    xw = (x.reshape(N * T, I).astype(np.float64) @ p['W_in'].astype(np.float64) + p['b']).astype('float32')
    ldh, ldw = (H + 7) // 8 * 8, 4 * H
    d_xw = G.dev(G.interleave_gates(xw, H))
    d_whid = G.dev(G.interleave_gates(p['W_hid'], H))
    wh, wl, sc = G.zeros((H, ldw), torch.float16), G.zeros((H, ldw), torch.float16), G.zeros((2,))
    G.call('ipavsr_f16_split', d_whid.data_ptr(), ldw, H, 4 * H, wh.data_ptr(), wl.data_ptr(), ldw, sc.data_ptr(),
           sc.data_ptr() + 4, 0, G.stream())
    d_peep, d_ci, d_hi, d_mask = G.dev(p['peep']), G.dev(p['cell_init']), G.dev(p['hid_init']), G.dev(mask)
    o_tc, o_ff = G.zeros((N * T, ldh)), G.zeros((N * T, ldh))
    G.call('ipavsr_lstm_fwd_f16', d_xw.data_ptr(), wh.data_ptr(), wl.data_ptr(), sc.data_ptr() + 4, ldw, d_peep.data_ptr(),
           d_ci.data_ptr(), d_hi.data_ptr(), d_mask.data_ptr(), o_tc.data_ptr(), None, None, None, N, T, H, ldh, 0, G.stream())
    nbytes = G.lib().ipavsr_lstm_workspace_bytes(N, T, H)
    ws = G.zeros(((nbytes + 3) // 4,))
    G.call('ipavsr_lstm_fwd', d_xw.data_ptr(), d_whid.data_ptr(), d_peep.data_ptr(), d_ci.data_ptr(), d_hi.data_ptr(),
           d_mask.data_ptr(), o_ff.data_ptr(), None, None, None, N, T, H, ldh, 0, 0, ws.data_ptr(), nbytes, G.stream())
    tc = G.host(o_tc)[:, :H].reshape(N, T, H)
    ff = G.host(o_ff)[:, :H].reshape(N, T, H)
    d_tc, d_ff, d_np = G.relerr(tc, ref64), G.relerr(ff, ref64), G.relerr(ref32, ref64)
    # growth of the drift with the step index: the signature of an expanding recurrence, not of a kernel error
    early = max(np.abs(tc[:, :5] - ref64[:, :5]).max(), 1e-12)
    late = np.abs(tc[:, -5:] - ref64[:, -5:]).max()
    print('LSTM drift, expanding recurrence (W_hid ~ N(0, 0.3), H=250, T=40): tensor-core %.2e, FFMA float32 %.2e, NumPy '
          'float32 %.2e of max|h|; tensor-core first 5 steps %.2e, last 5 steps %.2e' % (d_tc, d_ff, d_np, early, late))
    assert d_tc < 1e-3 and d_ff < 1e-3
    assert d_tc < 4 * max(d_ff, d_np) + 1e-6


@pytest.mark.parametrize('N,T,H,peep,backwards,sorted_lens', [(64, 20, 500, True, False, True), (64, 20, 500, False, True, True),
                                                               (48, 12, 320, True, True, False), (26, 40, 500, True, False, True),
                                                               (70, 9, 512, True, False, False)])
def test_lstm_steps_tensor_core(N, T, H, peep, backwards, sorted_lens):
    """ipavsr_lstm_{fwd,bwd}_f16_steps (wide layers: one tensor-core GEMM + one cell kernel per step) == the oracle, with and
    without the per-frame active-row counts of a length-sorted batch (tails of fewer than 4 rows take the exact product)."""
    rng = np.random.default_rng(N + T + H)
    lens = rng.integers(1, T + 1, size=N)
    lens[0] = T
    if sorted_lens:
        lens = np.sort(lens)[::-1].copy()
    I = 12
    p, x, mask = _lstm_inputs(rng, N, T, I, H, peep, lens)
    p['W_hid'] = (p['W_hid'] * 0.2).astype('float32')          # non-expanding recurrence (see the drift test)
    dout = rng.normal(size=(N, T, H)).astype('float32')
    out_ref, cache = ops.lstm_fwd(x, mask, p, backwards, np.float64)
    dx_ref, gr = ops.lstm_bwd(dout, cache, 5.0, np.float64)
    xw = (x.reshape(N * T, I).astype(np.float64) @ p['W_in'].astype(np.float64) + p['b']).astype('float32')
    ldh, ldw = (H + 7) // 8 * 8, 4 * H
    d_xw = G.dev(G.interleave_gates(xw, H))
    d_whid = G.dev(G.interleave_gates(p['W_hid'], H))
    wh, wl, sc = G.zeros((H, ldw), torch.float16), G.zeros((H, ldw), torch.float16), G.zeros((2,))
    G.call('ipavsr_f16_split', d_whid.data_ptr(), ldw, H, 4 * H, wh.data_ptr(), wl.data_ptr(), ldw, sc.data_ptr(),
           sc.data_ptr() + 4, 0, G.stream())
    d_peep = G.dev(p['peep']) if peep else None
    d_ci, d_hi, d_mask = G.dev(p['cell_init']), G.dev(p['hid_init']), G.dev(mask)
    nan = float('nan')
    d_out = torch.full((N * T, ldh), nan, dtype=torch.float32, device='cuda')
    d_gates = torch.full((N * T, 4 * H), nan, dtype=torch.float32, device='cuda')
    d_cell = torch.full((N * T, H), nan, dtype=torch.float32, device='cuda')
    d_hprev = torch.full((N * T, ldh), nan, dtype=torch.float32, device='cuda')
    assert G.lib().ipavsr_lstm_steps_supported(N, T, H, ldw)
    nbytes = G.lib().ipavsr_lstm_steps_workspace_bytes(N, T, H)
    ws = G.zeros(((nbytes + 3) // 4,))
    act = np.ascontiguousarray((lens[None, :] > np.arange(T)[:, None]).sum(1).astype(np.int32)) if sorted_lens else None
    import ctypes as C
    actp = act.ctypes.data_as(C.c_void_p) if act is not None else None
    G.call('ipavsr_lstm_fwd_f16_steps', d_xw.data_ptr(), d_whid.data_ptr(), wh.data_ptr(), wl.data_ptr(), sc.data_ptr() + 4, ldw,
           G.ptr(d_peep), d_ci.data_ptr(), d_hi.data_ptr(), d_mask.data_ptr(), d_out.data_ptr(), d_gates.data_ptr(),
           d_cell.data_ptr(), d_hprev.data_ptr(), N, T, H, ldh, int(backwards), actp, ws.data_ptr(), nbytes, G.stream())
    out = G.host(d_out)[:, :H].reshape(N, T, H)
    assert np.isfinite(out).all()
    assert G.relerr(out, out_ref) < 2e-5, G.relerr(out, out_ref)
    dop = np.zeros((N * T, ldh), 'float32')
    dop[:, :H] = dout.reshape(N * T, H)
    d_dout = G.dev(dop)
    d_dg = torch.full((N * T, 4 * H), nan, dtype=torch.float32, device='cuda')
    d_dpeep = G.zeros((3, H)) if peep else None
    d_dci, d_dhi = G.zeros((H,)), G.zeros((H,))
    dgh = torch.full((N * T, 4 * H), nan, dtype=torch.float16, device='cuda')
    dgl = torch.full((N * T, 4 * H), nan, dtype=torch.float16, device='cuda')
    dge = G.zeros((1,), torch.int32)
    G.call('ipavsr_lstm_bwd_f16_steps', d_dout.data_ptr(), d_whid.data_ptr(), wh.data_ptr(), wl.data_ptr(), sc.data_ptr() + 4,
           ldw, G.ptr(d_peep), d_ci.data_ptr(), d_mask.data_ptr(), d_gates.data_ptr(), d_cell.data_ptr(), d_dg.data_ptr(),
           G.ptr(d_dpeep), d_dci.data_ptr(), d_dhi.data_ptr(), N, T, H, ldh, int(backwards), 5.0, 0, dgh.data_ptr(),
           dgl.data_ptr(), dge.data_ptr(), actp, ws.data_ptr(), nbytes, G.stream())
    dgv = G.host(d_dg).astype(np.float64)
    assert np.isfinite(dgv).all()
    ex = int(G.host(dge)[0])
    rec = (G.host(dgh).astype(np.float64) + G.host(dgl).astype(np.float64) / G.F16_LO_SCALE) / 2.0 ** ex
    assert np.abs(rec - dgv).max() <= max(np.abs(dgv).max() * 2.0 ** -20, 1e-12)
    dG = G.deinterleave_gates(G.host(d_dg), H).astype(np.float64)
    tol = 3e-4
    assert G.relerr(dG.sum(0), gr['b']) < tol
    assert G.relerr((dG @ p['W_in'].astype(np.float64).T).reshape(N, T, I), dx_ref) < tol
    assert G.relerr(G.host(d_hprev)[:, :H].astype(np.float64).T @ dG, gr['W_hid']) < tol
    assert G.relerr(G.host(d_dci), gr['cell_init']) < tol, G.relerr(G.host(d_dci), gr['cell_init'])
    assert G.relerr(G.host(d_dhi), gr['hid_init']) < tol, G.relerr(G.host(d_dhi), gr['hid_init'])
    if peep:
        assert G.relerr(G.host(d_dpeep), gr['peep']) < tol
