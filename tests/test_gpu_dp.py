"""Two-GPU data-parallel training == single-GPU training on the global batch (NCCL all-reduce of the gradient arena,
device-side global loss normaliser, sync-BN).  Skipped on boxes with fewer than 2 GPUs."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def _train(name, rank, world, steps, q=None, port=None, dev_inputs=False):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, HERE)
    import model_util as MU
    from ipavsr_b200 import layers as L, parallel
    from ipavsr_b200.engine import get_engine
    from ipavsr_b200.function import function, tensor as T
    from ipavsr_b200.custom.objectives import temporal_softmax_loss, categorical_crossentropy
    from ipavsr_b200.custom.updates import nesterov_momentum
    if world > 1:
        os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                          LOCAL_RANK=str(rank))
        parallel.init_from_env('nccl')
    rng = np.random.default_rng(17)
    spec = MU.build(name, rng, C=6, H=16, win=3, fusiontype='concat')
    net = spec['net']
    MU.randomize_params(net, rng)
    N, Tn = 10, 9
    xs, mask, lens = MU.make_feed(rng, N, Tn, spec['dims'])
    y1 = rng.integers(0, 6, size=N).astype('int32')
    level = spec['level']
    y = y1 if level == 'seq' else np.repeat(y1[:, None], Tn, 1).astype('int32')
    ins = MU.input_layers(net)
    pred = L.get_output(net, deterministic=False)
    targets = T.imatrix('t') if level == 'frame' else T.ivector('t')
    cost = temporal_softmax_loss(pred, targets, ins['mask'].input_var) if level == 'frame' else \
        T.mean(categorical_crossentropy(pred, targets))
    params = L.get_all_params(net, trainable=True)
    order = [ins[n].input_var for n in spec['names']]
    train = function([order[0], targets, ins['mask'].input_var] + order[1:] + [T.iscalar('w')], cost,
                     updates=nesterov_momentum(cost, params, learning_rate=5e-2, momentum=0.9))
    if world > 1:
        parallel.attach(train.engine)
    lo, hi = parallel.shard_bounds(N, rank, world)
    losses = []
    # dev_inputs: everything resident in HBM — the call then returns as soon as the loss has been read back, which under
    # data parallelism is after the first gradient bucket's all-reduce (Engine._loss_early_copy)
    up = (lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()) if dev_inputs else (lambda a: a)
    for _ in range(steps):
        # device tensors for the mask so that the count is taken (and all-reduced) on the device
        losses.append(float(train(up(xs[0][lo:hi]), up(y[lo:hi]), torch.from_numpy(mask[lo:hi]).cuda(),
                                  *[up(x[lo:hi]) for x in xs[1:]], 3)))
    if dev_inputs and world > 1:
        assert train.engine.early_loss and train.engine._early_loss_ok
    vals = [p.get_value() for p in params]
    if q is not None:
        if rank == 0:
            q.put((losses, vals))
        torch.distributed.barrier()
        torch.distributed.destroy_process_group()
        return None
    return losses, vals


def _worker(rank, world, port, name, q, bucket):
    torch.cuda.set_device(rank)
    # a tiny bucket makes the overlapped all-reduce go out in many pieces during the backward walk; 0 = one all-reduce after
    # it; a negative bucket = the tiny bucket with device-resident inputs (early loss read-back)
    if bucket:
        os.environ['IPAVSR_AR_BUCKET'] = str(abs(bucket))
    else:
        os.environ['IPAVSR_AR_OVERLAP'] = '0'
    _train(name, rank, world, 3, q, port, dev_inputs=bucket < 0)


@pytest.mark.parametrize('bucket', [512, 0, -512])
@pytest.mark.parametrize('name', ['adenet_v2', 'adenet_v1'])
def test_two_gpu_training_matches_single_gpu(name, bucket):
    if torch.cuda.device_count() < 2:
        pytest.skip('needs 2 GPUs')
    ref_losses, ref_vals = _train(name, 0, 1, 3)
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    import zlib
    port = 29900 + (zlib.crc32(name.encode()) % 40) * 3 + (0 if not bucket else (1 if bucket > 0 else 2))
    procs = [ctx.Process(target=_worker, args=(r, 2, port, name, q, bucket)) for r in range(2)]
    for p in procs:
        p.start()
    losses, vals = q.get(timeout=300)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # SGD-type rule on purpose: Adam's g/sqrt(v) turns summation-order noise of near-zero gradients into O(lr) moves
    np.testing.assert_allclose(losses, ref_losses, rtol=2e-4)
    for a, b in zip(vals, ref_vals):
        assert np.abs(a - b).max() < 2e-5 * max(1.0, np.abs(b).max())
