"""Device-side batch assembly: `ipavsr_batch_gather` and the `utils/datagen.py` mirror against the
oracle (oracle/datagen.py) and the reference's own batches (tests/golden/datagen.npz)."""
import os

import numpy as np
import pytest
import torch

from oracle import datagen as OD
import gpu_util as G
import model_util as MU
from ipavsr_b200 import layers as L
from ipavsr_b200.function import function, tensor as T
from ipavsr_b200.utils import datagen as DG
from ipavsr_b200.custom.objectives import temporal_softmax_loss
from ipavsr_b200.custom.updates import adam

pytestmark = pytest.mark.gpu

GD = np.load(os.path.join(os.path.dirname(__file__), 'golden', 'datagen.npz'))


@pytest.mark.parametrize('U,F,N,pad', [(37, 50, 16, 0), (9, 1200, 7, 0), (21, 7, 33, 0), (12, 90, 5, 4), (5, 30, 1, 2),
                                       (300, 52, 512, 0)])
def test_batch_gather_kernel_matches_oracle(U, F, N, pad):
    rng = np.random.default_rng(U + F + N)
    seqlen = rng.integers(1, 41, size=U)
    seqlen[0], seqlen[-1] = 40, 1                  # a full-length and a one-frame utterance
    T = int(seqlen.max()) + (3 if pad else 0)
    total = int(seqlen.sum())
    ldd, ldx = F + pad, F + 2 * pad
    data = np.zeros((total, ldd), 'float32')
    data[:, :F] = rng.normal(size=(total, F))
    y = np.repeat(rng.integers(0, 26, size=U).astype('uint8'), seqlen)
    idxs = rng.integers(0, U, size=N)
    idxs[:2] = [0, U - 1][:len(idxs[:2])]
    integral = OD.compute_integral_len(seqlen)
    want_x, want_y, want_m = OD.lstm_batch(data[:, :F], y, seqlen, idxs, T)
    d_data, d_int = G.dev(data), torch.from_numpy(integral).cuda()
    d_len, d_idx = torch.from_numpy(seqlen.astype('int32')).cuda(), torch.from_numpy(idxs.astype('int32')).cuda()
    d_y = torch.from_numpy(y).cuda()
    sentinel = np.full((N * T, ldx), 7.5, 'float32')
    d_x = G.dev(sentinel)
    d_m, d_yb = torch.full((N, T), 9, dtype=torch.uint8, device='cuda'), torch.zeros(N, dtype=torch.uint8, device='cuda')
    G.call('ipavsr_batch_gather', d_data.data_ptr(), ldd, d_int.data_ptr(), d_len.data_ptr(), d_idx.data_ptr(),
           d_y.data_ptr(), d_x.data_ptr(), ldx, d_m.data_ptr(), d_yb.data_ptr(), N, T, F, G.stream())
    got = G.host(d_x)
    np.testing.assert_array_equal(got[:, :F].reshape(N, T, F), want_x)
    if ldx > F:
        np.testing.assert_array_equal(got[:, F:], sentinel[:, F:])       # the row pitch's padding is not touched
    np.testing.assert_array_equal(d_m.cpu().numpy(), want_m)
    np.testing.assert_array_equal(d_yb.cpu().numpy(), want_y)
    # mask / labels are optional
    d_x2 = G.zeros((N * T, ldx))
    G.call('ipavsr_batch_gather', d_data.data_ptr(), ldd, d_int.data_ptr(), d_len.data_ptr(), d_idx.data_ptr(), None,
           d_x2.data_ptr(), ldx, None, None, N, T, F, G.stream())
    np.testing.assert_array_equal(G.host(d_x2)[:, :F].reshape(N, T, F), want_x)


def test_python_api_reproduces_reference_batches():
    """Same names / arguments / shuffling as utils/datagen.py; batches are device tensors."""
    X, y, seqlen = GD['X'], GD['y'], GD['seqlen']
    integral = DG.compute_integral_len(seqlen)
    assert isinstance(integral, list) and integral == list(GD['integral'])
    for Tm in (20, 23):
        got = DG.gen_seq_batch_from_idx(X, GD['idxs'], seqlen, integral, Tm)
        assert got.is_cuda and got.dtype == torch.float32
        np.testing.assert_array_equal(got.cpu().numpy(), GD['seq_batch_T%d' % Tm])
    np.random.seed(1234)
    g = DG.gen_lstm_batch_random(X, y, seqlen, batchsize=4, shuffle=True)
    for k in range(5):
        xb, yb, mb, ib = next(g)
        np.testing.assert_array_equal(np.asarray(ib), GD['rand%d_idx' % k])
        np.testing.assert_array_equal(xb.cpu().numpy(), GD['rand%d_X' % k])
        np.testing.assert_array_equal(yb.cpu().numpy(), GD['rand%d_y' % k])
        np.testing.assert_array_equal(mb.cpu().numpy(), GD['rand%d_mask' % k])
    ds = DG.DeviceDataset(X, seqlen, y=y)
    g = DG.gen_lstm_batch_random(ds, None, None, batchsize=11, shuffle=False)
    for k in range(2):
        xb, yb, mb, ib = next(g)
        np.testing.assert_array_equal(xb.cpu().numpy(), GD['seq%d_X' % k])
        np.testing.assert_array_equal(mb.cpu().numpy(), GD['seq%d_mask' % k])
    with pytest.raises(ValueError):
        ds.gather([2], max_timesteps=5)            # utterance 2 has 20 frames
    with pytest.raises(IndexError):
        ds.gather([len(seqlen)])


def test_device_batches_feed_the_compiled_function():
    """A batch gathered on the device gives bit-identical outputs to the same batch assembled on the host."""
    rng = np.random.default_rng(3)
    spec = MU.build('adenet_v2', rng, fusiontype='concat')
    net = spec['net']
    ins = MU.input_layers(net)
    dims = spec['dims']
    U = 14
    seqlen = rng.integers(3, 12, size=U)
    total = int(seqlen.sum())
    streams = [rng.normal(size=(total, D)).astype('float32') for D in dims]
    y = np.repeat(rng.integers(0, 7, size=U).astype('uint8'), seqlen)
    idxs = rng.permutation(U)[:9]
    Tm = int(seqlen.max())
    window = T.iscalar('theta')
    val_fn = function([ins['input'].input_var, ins['mask'].input_var, ins['dct'].input_var, window],
                      L.get_output(net, deterministic=True))
    host = [OD.lstm_batch(s, y, seqlen, idxs, Tm) for s in streams]
    want = val_fn(host[0][0], host[0][2], host[1][0], 3)
    dsets = [DG.DeviceDataset(s, seqlen, y=y) for s in streams]
    xb0, mb = dsets[0].gather(idxs, Tm, with_mask=True)
    xb1 = dsets[1].gather(idxs, Tm)
    got = val_fn(xb0, mb, xb1, 3)
    np.testing.assert_array_equal(got, want)


def test_training_from_the_device_generator_matches_host_batches():
    """The runner loop (runners/2stream_dct.py:312-326) with gen_lstm_batch_random on the device: same costs, step by
    step, as the reference generator's host batches under the same NumPy seed."""
    rng = np.random.default_rng(5)
    U = 17
    seqlen = rng.integers(3, 12, size=U)
    total = int(seqlen.sum())
    dims = MU.build('adenet_v2', np.random.default_rng(21), fusiontype='sum')['dims']
    streams = [rng.normal(size=(total, D)).astype('float32') for D in dims]
    y = np.repeat(rng.integers(0, 7, size=U).astype('uint8'), seqlen)
    integral = OD.compute_integral_len(seqlen)
    Tm = int(seqlen.max())
    costs = {}
    for device_feed in (False, True):
        spec = MU.build('adenet_v2', np.random.default_rng(21), fusiontype='sum')
        net = spec['net']
        ins = MU.input_layers(net)
        targets, window = T.imatrix('targets'), T.iscalar('theta')
        cost = temporal_softmax_loss(L.get_output(net, deterministic=False), targets, ins['mask'].input_var)
        train = function([ins['input'].input_var, targets, ins['mask'].input_var, ins['dct'].input_var, window], cost,
                         updates=adam(cost, L.get_all_params(net, trainable=True), learning_rate=1e-3))
        np.random.seed(3)
        out = []
        if device_feed:
            ds0, ds1 = DG.DeviceDataset(streams[0], seqlen, y=y), DG.DeviceDataset(streams[1], seqlen)
            gen = DG.gen_lstm_batch_random(ds0, None, None, batchsize=6)
            for _ in range(4):
                X, yb, m, idx = next(gen)
                yy = yb.reshape(-1, 1).expand(-1, m.shape[-1])
                X2 = DG.gen_seq_batch_from_idx(ds1, idx, seqlen, integral, Tm)
                out.append(float(train(X, yy, m, X2, 3)))
        else:
            sched = OD.batch_schedule(U, 6)
            for _ in range(4):
                idx = next(sched)
                X, yb, m = OD.lstm_batch(streams[0], y, seqlen, idx)
                yy = yb.reshape((-1, 1)).repeat(m.shape[-1], axis=-1)
                X2 = OD.seq_batch_from_idx(streams[1], idx, seqlen, integral, Tm)
                out.append(float(train(X, yy, m, X2, 3)))
        costs[device_feed] = out
    # same batches, same initial weights; split-K GEMMs accumulate with atomics, so two runs agree to rounding, not bit for bit
    np.testing.assert_allclose(costs[True], costs[False], rtol=2e-5)
