"""world_size-2 gloo tests (CPU) of the data-parallel algebra: sharding utterances, normalising by the GLOBAL
count and summing gradients across ranks reproduces the single-process gradient of the global batch — for the
frame-level loss (global mask sum), the sequence-level loss (global N) and with BatchNorm statistics all-reduced.
The arithmetic is the NumPy oracle; the collective plumbing is the product's (ipavsr_b200/parallel.py)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def _worker(rank, world, port, name, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, HERE)
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    from ipavsr_b200 import parallel, layers as L
    from oracle.net import OracleNet
    from oracle import ops
    import model_util as MU
    parallel.init_from_env('gloo')
    rng = np.random.default_rng(5)
    spec = MU.build(name, rng, C=5, H=6, win=2, fusiontype='concat')
    net = spec['net']
    MU.randomize_params(net, rng)
    N, T = 7, 6                       # odd N: uneven shards
    xs, mask, lens = MU.make_feed(rng, N, T, spec['dims'])
    y1 = rng.integers(0, 5, size=N)
    level = spec['level']
    y = y1 if level == 'seq' else np.repeat(y1[:, None], T, 1)
    lo, hi = parallel.shard_bounds(N, rank, world)
    feed = {k: v[lo:hi] for k, v in zip(spec['names'], xs)}
    feed['mask'] = mask[lo:hi]
    o = OracleNet(net, np.float64)
    out = o.forward(feed, 2, deterministic=True)
    if level == 'frame':
        count = parallel.global_normaliser(mask[lo:hi].sum())
        n = (hi - lo) * T
        q_ = ops.softmax_rows(out.reshape(n, -1))
        yy, mm = y[lo:hi].reshape(n), mask[lo:hi].reshape(n).astype(np.float64)
        loss_local = -(mm * np.log(q_[np.arange(n), yy])).sum() / count
        dq = q_.copy()
        dq[np.arange(n), yy] -= 1
        dout = (dq * (mm / count)[:, None]).reshape(out.shape)
    else:
        count = parallel.global_normaliser(hi - lo)
        n = hi - lo
        loss_local = -np.log(out[np.arange(n), y[lo:hi]]).sum() / count
        dout = np.zeros_like(out)
        dout[np.arange(n), y[lo:hi]] = -1.0 / (out[np.arange(n), y[lo:hi]] * count)
    grads = o.backward(dout)
    params = L.get_all_params(net, trainable=True)
    summed = parallel.allreduce_host_grads([grads[p] for p in params])
    loss = parallel.global_normaliser(loss_local)
    if rank == 0:
        full = OracleNet(net, np.float64)
        ffeed = dict(zip(spec['names'], xs))
        ffeed['mask'] = mask
        ref_loss, _, ref_grads = full.loss_and_grads(ffeed, 2, y, mask, 'categorical_crossentropy' if level == 'seq'
                                                     else 'temporal_softmax', deterministic=True)
        err = max(float(np.abs(a - b.reshape(a.shape)).max() / max(np.abs(b).max(), 1e-12))
                  for a, b in zip(summed, ref_grads))
        q.put((abs(loss - float(ref_loss)), err, count))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize('name', ['adenet_v2', 'deltanet', 'adenet_4stream'])
def test_sharded_gradients_equal_global_gradients(name):
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = 29500 + (abs(hash(name)) % 400)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, name, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = q.get(timeout=180)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    dloss, gerr, count = res
    assert dloss < 1e-12 and gerr < 1e-10, res


def test_shard_bounds_cover_everything():
    from ipavsr_b200.parallel import shard_bounds
    for n in (0, 1, 7, 26, 512, 4096):
        for w in (1, 2, 3, 8):
            spans = [shard_bounds(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def _ae_worker(rank, world, port, q):
    """Auto-encoder fine-tuning objective under data parallelism, the engine's rule: every rank's squared error is
    normalised by the GLOBAL element count, rank 0 alone adds the L2 penalty (value and 2·l2·W), gradients and loss are summed."""
    sys.path.insert(0, ROOT)
    sys.path.insert(0, HERE)
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    from ipavsr_b200 import parallel, layers as L, nonlinearities as nl
    from ipavsr_b200.function import tensor as T
    from oracle.net import OracleNet
    parallel.init_from_env('gloo')
    rng = np.random.default_rng(9)
    sizes = (10, 7, 3, 7, 10)
    l_in = L.InputLayer((None, None, sizes[0]), T.tensor3('x'), name='input')
    l = L.ReshapeLayer(l_in, (-1, sizes[0]))
    for i in range(4):
        l = L.DenseLayer(l, sizes[i + 1], W=rng.normal(size=(sizes[i], sizes[i + 1])).astype('float32'),
                         b=rng.normal(size=sizes[i + 1]).astype('float32'),
                         nonlinearity=nl.linear if i in (1, 3) else nl.sigmoid, name='l%d' % (i + 1))
    M, l2c = 13, 0.005                                        # odd row count: uneven shards
    X = rng.normal(size=(M, 1, sizes[0]))
    lo, hi = parallel.shard_bounds(M, rank, world)
    o = OracleNet(l, np.float64)
    out = o.forward({'input': X[lo:hi]}, 0, deterministic=False)
    count = parallel.global_normaliser((hi - lo) * sizes[0])
    d = out - X[lo:hi].reshape(out.shape)
    loss_local = (d * d).sum() / count
    grads = o.backward(2.0 * d / count)
    params = L.get_all_params(l, trainable=True)
    if rank == 0:
        for p in L.get_all_params(l, regularizable=True):
            w = p.get_value().astype(np.float64)
            loss_local += l2c * (w * w).sum()
            grads[p] = grads[p] + 2 * l2c * w
    summed = parallel.allreduce_host_grads([grads[p] for p in params])
    loss = parallel.global_normaliser(loss_local)
    if rank == 0:
        ref_loss, _, ref_grads = OracleNet(l, np.float64).loss_and_grads({'input': X}, 0, X.reshape(M, -1), None,
                                                                         'squared_error', l2=l2c)
        err = max(float(np.abs(a - b.reshape(a.shape)).max() / max(np.abs(b).max(), 1e-12)) for a, b in zip(summed, ref_grads))
        q.put((abs(loss - float(ref_loss)), err, count))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_autoencoder_objective_equals_the_global_one():
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    procs = [ctx.Process(target=_ae_worker, args=(r, 2, 29931, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = q.get(timeout=180)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    dloss, gerr, count = res
    assert count == 130 and dloss < 1e-12 and gerr < 1e-10, res
