"""Uninitialised-read detector for the engine: poisons the caching allocator's free blocks with NaN, then runs one
forward/backward of a builder and prints the per-parameter gradient error against the oracle (NaN = a buffer was read
before it was written).   python tools/debug_poison.py adenet_v2_1:concat adenet_v2_2:concat"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np
import torch
import model_util as MU
import test_gpu_models as TG
from ipavsr_b200 import layers as L
from ipavsr_b200.engine import Engine


def poison():
    big = torch.full((1 << 28,), float('nan'), device='cuda')
    small = [torch.full((n,), float('nan'), device='cuda') for n in (1 << 10, 1 << 12, 1 << 14, 1 << 16, 1 << 18) for _ in range(200)]
    torch.cuda.synchronize()
    del big, small


for arg in sys.argv[1:]:
    name, fus = arg.split(':')
    for seed in (TG._seed(name), 1, 2):
        spec, net, feed, mask, y, dm, win = TG._case(name, seed, fus)
        loss_ref, out_ref, grads_ref = TG._oracle(net, feed, win, y, mask, spec['level'], dm)
        poison()
        eng = Engine(net, gemm_mode='fp32')
        ins = MU.input_layers(net)
        run, out = eng.forward({ins[k]: v for k, v in feed.items()}, win, deterministic=False, train=True, dropout_masks=dm,
                               update_bn=False)
        probs = eng.read(out).reshape(out_ref.shape)
        eng.loss_and_backward(run, out, 'categorical_crossentropy' if spec['level'] == 'seq' else 'temporal_softmax', y, mask,
                              count=float(mask.sum()))
        params = L.get_all_params(net, trainable=True)
        grads = eng.param_grads(params)
        gmax = max(np.abs(g).max() for g in grads_ref)
        bad = []
        for p, g, gr in zip(params, grads, grads_ref):
            e = np.abs(g - gr).max() / max(np.abs(gr).max(), 2e-2 * gmax)
            if not e < 2e-3:
                bad.append('%s %.3g' % (p.name, e))
        print('%s %s seed %d: probs err %.2e, loss %.6f vs %.6f, bad grads: %s' % (
            name, fus, seed, np.abs(probs - out_ref).max(), eng.read_loss(), loss_ref, bad or 'none'), flush=True)
