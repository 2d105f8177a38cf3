"""Packed / length-sorted execution (engine._PackPlan, csrc/pack.cu) and the tile-wide frame skipping of the tensor-core
LSTM kernels: the row movers against NumPy indexing, the recurrence kernels on masks with tile-wide masked frames against
the oracle, and whole networks run packed against the same networks run in the reference's padded layout and against
the float64 oracle.  Run with -m gpu on the B200."""
import numpy as np
import pytest
import torch

from ipavsr_b200 import layers as L
from ipavsr_b200.engine import Engine, _PackPlan
from oracle import ops
from oracle.net import OracleNet
import gpu_util as G
import model_util as MU
from test_gpu_kernels import _lstm_inputs

pytestmark = pytest.mark.gpu


# ---------------------------------------------------------------------------------------------------------
# row movers
# ---------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize('rows_in,F,ld_in,ld_out', [(100, 1200, 1200, 1200), (57, 90, 90, 96), (33, 50, 56, 56),
                                                    (9, 7, 7, 9), (300, 26, 32, 26)])
@pytest.mark.parametrize('host_src', [False, True])
def test_gather_rows(rows_in, F, ld_in, ld_out, host_src):
    rng = np.random.default_rng(rows_in + F)
    src = rng.normal(size=(rows_in, ld_in)).astype('float32')
    idx = rng.integers(-1, rows_in, size=2 * rows_in + 3).astype('int32')
    idx[0] = -1
    fill = rng.normal(size=(ld_in,)).astype('float32')
    d_idx = G.dev(idx)
    for use_fill in (False, True):
        if host_src:
            s = torch.from_numpy(src).pin_memory()      # pinned host memory is read by the kernel itself
        else:
            s = G.dev(src)
        d_fill = G.dev(fill) if use_fill else None
        dst = torch.full((len(idx), ld_out), 7.0, dtype=torch.float32, device='cuda')
        G.call('ipavsr_gather_rows', s.data_ptr(), 4 * ld_in, dst.data_ptr(), 4 * ld_out, 4 * F, d_idx.data_ptr(),
               G.ptr(d_fill), len(idx), G.stream())
        got = G.host(dst)
        want = np.where(idx[:, None] >= 0, src[np.maximum(idx, 0), :F], fill[None, :F] if use_fill else 0.0)
        np.testing.assert_array_equal(got[:, :F], want)
        assert (got[:, F:] == 7.0).all()               # columns beyond the row stay untouched


def test_gather_rows_bytes():
    """uint8 / int32 rows (masks, targets)."""
    rng = np.random.default_rng(3)
    m = rng.integers(0, 2, size=(40, 13)).astype('uint8')
    idx = rng.permutation(40).astype('int32')
    d_m, d_idx = G.dev(m), G.dev(idx)
    out = torch.zeros(40, 13, dtype=torch.uint8, device='cuda')
    G.call('ipavsr_gather_rows', d_m.data_ptr(), 13, out.data_ptr(), 13, 13, d_idx.data_ptr(), None, 40, G.stream())
    np.testing.assert_array_equal(G.host(out), m[idx])


@pytest.mark.parametrize('M,N,ld', [(1000, 50, 56), (38400, 50, 56), (7, 3, 3), (513, 130, 136)])
def test_colsum_masked(M, N, ld):
    rng = np.random.default_rng(M + N)
    x = rng.normal(size=(M, ld)).astype('float32')
    mask = rng.integers(0, 2, size=M).astype('uint8')
    d_x, d_m = G.dev(x), G.dev(mask)
    for invert in (0, 1):
        out = torch.full((N,), 3.0, dtype=torch.float32, device='cuda')
        G.call('ipavsr_colsum_masked', d_x.data_ptr(), ld, d_m.data_ptr(), invert, out.data_ptr(), M, N, 0, G.stream())
        want = x[(mask != 0) != bool(invert), :N].astype(np.float64).sum(0)
        assert np.abs(G.host(out) - want).max() < 1e-5 * max(1.0, np.abs(want).max()) * np.sqrt(M)
        G.call('ipavsr_colsum_masked', d_x.data_ptr(), ld, d_m.data_ptr(), invert, out.data_ptr(), M, N, 1, G.stream())
        assert np.abs(G.host(out) - 2 * want).max() < 2e-5 * max(1.0, np.abs(want).max()) * np.sqrt(M)


def test_pack_plan_tables():
    rng = np.random.default_rng(0)
    N, T = 37, 11
    lens = rng.integers(0, T + 1, size=N)
    lens[3] = T
    plan = _PackPlan(lens, T)
    t = dict(plan.tables)
    x = rng.normal(size=(N * T, 3))
    mask = (np.arange(T)[None, :] < lens[:, None])
    x[~mask.reshape(-1)] = 0
    packed = np.where(t['pack'][:, None] >= 0, x[np.maximum(t['pack'], 0)], 0.0)
    assert plan.M == lens.sum() and len(packed) == plan.M + 1 and (packed[-1] == 0).all()
    # lengths are non-increasing in the sorted order, ties keep their original order
    assert (np.diff(plan.lens_sorted) <= 0).all()
    srt = x.reshape(N, T, 3)[plan.order_host].reshape(N * T, 3)
    np.testing.assert_array_equal(packed[t['unpack']], srt)                # unpack(pack(x)) = x in sorted order
    np.testing.assert_array_equal(srt[t['valid']], packed[:-1])
    np.testing.assert_array_equal(x[t['perm']], srt)
    np.testing.assert_array_equal(srt[t['unperm']], x)
    np.testing.assert_array_equal(t['order'][t['inv']], np.arange(N))
    np.testing.assert_array_equal(t['mask'].view(np.uint8)[:N * T].reshape(N, T), mask[plan.order_host])


# ---------------------------------------------------------------------------------------------------------
# tensor-core LSTM kernels: frames at which a whole 32-utterance tile is masked are skipped
# ---------------------------------------------------------------------------------------------------------
def _masks(kind, rng, N, T):
    if kind == 'sorted':                 # what the packed engine produces: later tiles are shorter
        lens = np.sort(rng.integers(1, T + 1, size=N))[::-1].copy()
        lens[0] = T
        return (np.arange(T)[None, :] < lens[:, None]).astype('uint8')
    if kind == 'short':                  # every utterance much shorter than T
        lens = rng.integers(1, max(2, T // 3), size=N)
        return (np.arange(T)[None, :] < lens[:, None]).astype('uint8')
    if kind == 'gaps':                   # not a prefix mask: frames masked for everybody in the middle and at the start
        m = (rng.random((N, T)) < 0.8).astype('uint8')
        m[:, 0] = 0
        m[:, T // 2: T // 2 + 2] = 0
        return m
    if kind == 'empty_tile':             # one whole tile of zero-length utterances
        lens = rng.integers(1, T + 1, size=N)
        lens[32:64] = 0
        return (np.arange(T)[None, :] < lens[:, None]).astype('uint8')
    raise KeyError(kind)


@pytest.mark.parametrize('kind', ['sorted', 'short', 'gaps', 'empty_tile'])
@pytest.mark.parametrize('N,T,H,peep,backwards', [(100, 40, 250, True, False), (100, 40, 250, False, True),
                                                    (70, 13, 40, True, True), (70, 13, 40, True, False)])
def test_lstm_tensor_core_skips_masked_frames(kind, N, T, H, peep, backwards):
    rng = np.random.default_rng(N + T + H + len(kind))
    I = 12
    p, x, _ = _lstm_inputs(rng, N, T, I, H, peep, np.full(N, T))
    mask = _masks(kind, rng, N, T)
    p['W_hid'] = (p['W_hid'] * (0.25 if H >= 100 else 1.0)).astype('float32')
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
    nan = float('nan')                     # the kernel must write every element it is responsible for
    d_out = torch.full((N * T, ldh), nan, dtype=torch.float32, device='cuda')
    d_gates = torch.full((N * T, 4 * H), nan, dtype=torch.float32, device='cuda')
    d_cell = torch.full((N * T, H), nan, dtype=torch.float32, device='cuda')
    d_hprev = torch.full((N * T, ldh), nan, dtype=torch.float32, device='cuda')
    G.call('ipavsr_lstm_fwd_f16', d_xw.data_ptr(), wh.data_ptr(), wl.data_ptr(), sc.data_ptr() + 4, ldw, G.ptr(d_peep),
           d_ci.data_ptr(), d_hi.data_ptr(), d_mask.data_ptr(), d_out.data_ptr(), d_gates.data_ptr(), d_cell.data_ptr(),
           d_hprev.data_ptr(), N, T, H, ldh, int(backwards), G.stream())
    out = G.host(d_out)[:, :H].reshape(N, T, H)
    assert np.isfinite(out).all()
    assert G.relerr(out, out_ref) < 2e-5, G.relerr(out, out_ref)
    for a in (d_gates, d_cell, d_hprev[:, :H]):
        assert np.isfinite(G.host(a)).all()
    # hprev feeds dW_hid = hprev^T dgates: compare that product with the oracle's gradient below
    dop = np.zeros((N * T, ldh), 'float32')
    dop[:, :H] = dout.reshape(N * T, H)
    d_dout = G.dev(dop)
    d_dg = torch.full((N * T, 4 * H), nan, dtype=torch.float32, device='cuda')
    d_dpeep = G.zeros((3, H)) if peep else None
    d_dci, d_dhi, d_db = G.zeros((H,)), G.zeros((H,)), G.zeros((4 * H,))
    dgh = torch.full((N * T, 4 * H), nan, dtype=torch.float16, device='cuda')
    dgl = torch.full((N * T, 4 * H), nan, dtype=torch.float16, device='cuda')
    dge = G.zeros((1,), torch.int32)
    nbytes = G.lib().ipavsr_lstm_workspace_bytes(N, T, H)
    ws = G.zeros(((nbytes + 3) // 4,))
    G.call('ipavsr_lstm_bwd_f16', d_dout.data_ptr(), d_whid.data_ptr(), wh.data_ptr(), wl.data_ptr(), sc.data_ptr() + 4,
           ldw, G.ptr(d_peep), d_ci.data_ptr(), d_mask.data_ptr(), d_gates.data_ptr(), d_cell.data_ptr(),
           d_dg.data_ptr(), G.ptr(d_dpeep), d_dci.data_ptr(), d_dhi.data_ptr(), N, T, H, ldh, int(backwards), 5.0, 0,
           d_db.data_ptr(), dgh.data_ptr(), dgl.data_ptr(), dge.data_ptr(), ws.data_ptr(), nbytes, G.stream())
    dgv = G.host(d_dg).astype(np.float64)
    assert np.isfinite(dgv).all()
    ex = int(G.host(dge)[0])
    rec = (G.host(dgh).astype(np.float64) + G.host(dgl).astype(np.float64) / G.F16_LO_SCALE) / 2.0 ** ex
    assert np.abs(rec - dgv).max() <= max(np.abs(dgv).max() * 2.0 ** -20, 1e-12)
    dG = G.deinterleave_gates(G.host(d_dg), H).astype(np.float64)
    tol = 3e-4
    assert G.relerr(G.host(d_db), G.host(d_dg).astype(np.float64).sum(0)) < 1e-5
    assert G.relerr(dG.sum(0), gr['b']) < tol
    assert G.relerr((dG @ p['W_in'].astype(np.float64).T).reshape(N, T, I), dx_ref) < tol
    assert G.relerr(G.host(d_hprev)[:, :H].astype(np.float64).T @ dG, gr['W_hid']) < tol
    assert G.relerr(G.host(d_dci), gr['cell_init']) < tol, G.relerr(G.host(d_dci), gr['cell_init'])
    assert G.relerr(G.host(d_dhi), gr['hid_init']) < tol, G.relerr(G.host(d_dhi), gr['hid_init'])
    if peep:
        assert G.relerr(G.host(d_dpeep), gr['peep']) < tol


# ---------------------------------------------------------------------------------------------------------
# whole networks: packed / length-sorted run == padded run == oracle
# ---------------------------------------------------------------------------------------------------------
def _case(name, seed, fusiontype, N, T, H=12, C=7, win=3):
    rng = np.random.default_rng(seed)
    spec = MU.build(name, rng, C=C, H=H, win=win, fusiontype=fusiontype)
    net = spec['net']
    MU.randomize_params(net, rng)
    lens = rng.integers(1, T + 1, size=N)
    lens[N // 2] = T
    xs, mask, lens = MU.make_feed(rng, N, T, spec['dims'], lens=lens)
    y1 = rng.integers(0, C, size=N).astype('int32')
    y = y1 if spec['level'] == 'seq' else np.repeat(y1[:, None], T, 1).astype('int32')
    feed = dict(zip(spec['names'], xs))
    feed['mask'] = mask
    dm = MU.dropout_masks_for(net, rng, N, T)
    return spec, net, feed, mask, y, dm, win


def _run(net, feed, win, y, mask, level, dm, mode, packed, device_inputs=False):
    eng = Engine(net, gemm_mode=mode, packed=packed)
    ins = MU.input_layers(net)
    if device_inputs:
        dfeed = {ins[k]: torch.from_numpy(np.ascontiguousarray(v)).cuda() for k, v in feed.items()}
    else:
        dfeed = {ins[k]: v for k, v in feed.items()}
    run, out = eng.forward(dfeed, win, deterministic=False, train=True, dropout_masks=dm, update_bn=False)
    assert (run.plan is not None) == (packed == 'force')
    probs = eng.read(out)
    loss_name = 'categorical_crossentropy' if level == 'seq' else 'temporal_softmax'
    yy = torch.from_numpy(y).cuda() if device_inputs else y
    eng.loss_and_backward(run, out, loss_name, yy, dfeed[ins['mask']], count=float(mask.sum()))
    params = L.get_all_params(net, trainable=True)
    return probs, float(eng.read_loss()), eng.param_grads(params), params, run


@pytest.mark.parametrize('name,fusiontype,device_inputs',
                         [('adenet_v2', 'concat', False), ('adenet_v2', 'adasum', True), ('adenet_3stream', 'concat', True),
                          ('adenet_v1', 'sum', False), ('adenet_v3', 'sum', False), ('deltanet', 'sum', True),
                          ('adenet_4stream', 'concat', False), ('lstm_classifier_baseline', 'sum', False),
                          ('deltanet_v1', 'sum', True)])
@pytest.mark.parametrize('mode', ['fp32', 'f16x3'])
def test_packed_run_matches_padded_run_and_oracle(name, fusiontype, device_inputs, mode):
    spec, net, feed, mask, y, dm, win = _case(name, 21, fusiontype, N=41, T=13)
    level = spec['level']
    p0, l0, g0, params, _ = _run(net, feed, win, y, mask, level, dm, mode, 'off')
    p1, l1, g1, _, run1 = _run(net, feed, win, y, mask, level, dm, mode, 'force', device_inputs)
    # against the float64 oracle (original utterance order); rectify branches at rounding distance from 0 follow the
    # device (model_util.rectify_aligner)
    lname = 'categorical_crossentropy' if level == 'seq' else 'temporal_softmax'
    align, flips = MU.rectify_aligner(net, run1, 41, 13)
    loss_ref, out_ref, grads_ref = OracleNet(net, np.float64).loss_and_grads(
        feed, win, y, mask, lname, deterministic=False, dropout_masks=dm, update_bn=False, after_forward=align)
    pr = p1.reshape(out_ref.shape)
    assert np.abs(pr - out_ref).max() / np.abs(out_ref).max() < 1e-4
    assert (pr.argmax(-1) == out_ref.argmax(-1)).all()
    assert abs(l1 - loss_ref) < 1e-4 * abs(loss_ref)
    for p, err in MU.grad_errors(params, g1, grads_ref):
        assert err < 1e-4, (p.name, err, flips)
    # the two layouts differ only in summation order (and in the utterances an LSTM tile groups together); a rectify unit at
    # rounding distance from 0 may still take different branches in the two runs, so the direct comparison is made when
    # the oracle saw no such unit
    assert np.abs(p1 - p0).max() < 2e-5 * np.abs(p0).max(), np.abs(p1 - p0).max()
    assert abs(l1 - l0) < 2e-5 * abs(l0)
    if not flips:
        for p, err in MU.grad_errors(params, g1, g0):
            assert err < 1e-4, (p.name, err)


@pytest.mark.parametrize('upload', ['dma', 'gather'])
def test_packed_pinned_host_inputs_and_prefetch(upload, monkeypatch):
    """Pinned host streams are uploaded ragged — one copy-engine transfer per utterance ('dma', the default) or the gather
    kernel reading host memory ('gather'); prefetch (immediate and deferred) stages the same plan; pageable inputs take the
    copy-then-gather route.  All give the same probabilities."""
    monkeypatch.setenv('IPAVSR_HOST_UPLOAD', upload)
    from ipavsr_b200.function import function, tensor as T
    spec, net, feed, mask, y, dm, win = _case('adenet_v2', 5, 'concat', N=64, T=20)
    ins = MU.input_layers(net)
    val = function([ins['input'].input_var, ins['mask'].input_var, ins['dct'].input_var, T.iscalar('w')],
                   L.get_output(net, deterministic=True), packed='force')
    ref = OracleNet(net, np.float64).forward(feed, win, deterministic=True)
    a = val(feed['input'], mask, feed['dct'], win)
    pin = lambda v: torch.from_numpy(np.ascontiguousarray(v)).pin_memory()
    hx, hm, hd = pin(feed['input']), pin(mask), pin(feed['dct'])
    b = val(hx, hm, hd, win)
    val.prefetch(hx, hm, hd, win)
    assert len(val.engine._prefetched) == 1
    c = val(hx, hm, hd, win)
    assert len(val.engine._prefetched) == 0
    val.prefetch(hx, hm, hd, win, defer=True)          # staged by the next call, behind its own kernels
    assert len(val.engine._prefetched) == 0
    c2 = val(hx, hm, hd, win)
    assert len(val.engine._prefetched) == 1
    c3 = val(hx, hm, hd, win)
    assert len(val.engine._prefetched) == 0
    np.testing.assert_array_equal(a, b)
    np.testing.assert_array_equal(a, c)
    np.testing.assert_array_equal(a, c2)
    np.testing.assert_array_equal(a, c3)
    assert np.abs(a - ref).max() / np.abs(ref).max() < 1e-4
    # a mask that is not a prefix mask falls back to the padded layout (and still matches the oracle)
    m2 = mask.copy()
    m2[3, 0] = 0
    ref2 = OracleNet(net, np.float64).forward(dict(feed, mask=m2), win, deterministic=True)
    got2 = val(feed['input'], m2, feed['dct'], win)
    assert np.abs(got2 - ref2).max() / np.abs(ref2).max() < 1e-4


# ---------------------------------------------------------------------------------------------------------
# derived streams: diff images and DCT(+deltas) computed on the device from the raw stream's packed frames
# ---------------------------------------------------------------------------------------------------------
def _derived_host(raw, lens, image_shape, K):
    """The reference's host pipeline on the valid frames, padded back: compute_diff_images (runners/3stream.py:95-96);
    compute_dct_features + concat_first_second_deltas, float32 (avletters/preprocess_images.py:20-21, bimodal.py:351)."""
    from oracle import preprocessing as OP
    N, T_, D = raw.shape
    packed = np.concatenate([raw[i, :lens[i]] for i in range(N)], 0)
    diff_p = OP.compute_diff_images(packed, lens)
    dct_p = OP.concat_first_second_deltas(OP.compute_dct_features(packed, image_shape, K, 'zigzag'), lens).astype('float32')
    diff = np.zeros((N, T_, D), 'float32')
    dct = np.zeros((N, T_, 3 * K), 'float32')
    o = 0
    for i in range(N):
        diff[i, :lens[i]] = diff_p[o:o + lens[i]]
        dct[i, :lens[i]] = dct_p[o:o + lens[i]]
        o += lens[i]
    return diff, dct


@pytest.mark.parametrize('device_raw', [False, True])
def test_derived_streams_match_host_preprocessing(device_raw):
    from ipavsr_b200 import modelzoo, init
    from ipavsr_b200.function import function, tensor as T
    from ipavsr_b200.derived import DiffImages, DctFeatures
    rng = np.random.default_rng(31)
    np.random.seed(31)
    ish, K, H, C, win, N, T_ = (6, 8), 5, 12, 7, 3, 40, 14
    D = ish[0] * ish[1]
    aes = [MU.ae_tuple(rng, d) for d in (D, D, 3 * K)]
    v = [T.tensor3('s%d' % i) for i in range(3)]
    m = T.matrix('mask', dtype='uint8')
    net, _ = modelzoo.adenet_3stream.create_model(aes[0], aes[1], aes[2], (None, None, D), v[0], (None, None, D), v[1],
                                                  (None, None, 3 * K), v[2], (None, None), m, H, win, C, 'concat',
                                                  init.Orthogonal(), True)
    MU.randomize_params(net, rng)
    lens = rng.integers(2, T_ + 1, size=N)
    lens[5] = T_
    (raw,), mask, _ = MU.make_feed(rng, N, T_, [D], lens=lens)
    diff, dct = _derived_host(raw, lens, ish, K)
    val = function([v[0], v[1], v[2], m, T.iscalar('w')], L.get_output(net, deterministic=True))
    want = val(raw, diff, dct, mask, win)
    ref = OracleNet(net, np.float64).forward({'s1_im': raw, 's2_im': diff, 's3_im': dct, 'mask': mask}, win,
                                             deterministic=True)
    assert np.abs(want - ref).max() / np.abs(ref).max() < 1e-4
    r = torch.from_numpy(raw).cuda() if device_raw else torch.from_numpy(raw).pin_memory()
    got = val(r, DiffImages(r), DctFeatures(r, ish, K), mask, win)
    # the diff images are bit-exact, the DCT projection carries float32 accumulation error (2e-5 of the largest coefficient)
    assert np.abs(got - want).max() / np.abs(want).max() < 2e-4, np.abs(got - want).max()
    assert (got.argmax(-1) == ref.argmax(-1)).all()
    # prefetch with derived streams: only the raw stream is staged
    if not device_raw:
        args = (r, DiffImages(r), DctFeatures(r, ish, K), torch.from_numpy(mask).pin_memory(), win)
        val.prefetch(*args)
        got2 = val(*args)
        np.testing.assert_array_equal(got2, got)
    # training through derived streams == training on the host-computed streams
    from ipavsr_b200.custom.objectives import temporal_softmax_loss
    from ipavsr_b200.custom.updates import sgd
    tg = T.imatrix('t')
    cost = temporal_softmax_loss(L.get_output(net, deterministic=False), tg, m)
    params = L.get_all_params(net, trainable=True)
    train = function([v[0], v[1], v[2], tg, m, T.iscalar('w')], cost, updates=sgd(cost, params, learning_rate=0.0))
    y = np.repeat(rng.integers(0, C, size=(N, 1)), T_, 1).astype('int32')
    la = train(raw, diff, dct, y, mask, win)
    ga = train.engine.param_grads(params)
    lb = train(r, DiffImages(r), DctFeatures(r, ish, K), y, mask, win)
    gb = train.engine.param_grads(params)
    assert abs(la - lb) < 1e-4 * abs(la)
    for p, err in MU.grad_errors(params, gb, ga):
        assert err < 1e-3, (p.name, err)          # the derived DCT projection carries float32 accumulation error (2e-5)
