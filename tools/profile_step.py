"""Per-entry-point device time of one bench training step (CUDA events around each C-ABI call).
    python tools/profile_step.py [--mode tf32x3] [--batch 960]"""
import argparse, os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import bench
from ipavsr_b200 import layers as L, _lib, engine as E
from ipavsr_b200.function import function, tensor as T
from ipavsr_b200.custom.objectives import temporal_softmax_loss
from ipavsr_b200.custom.updates import adam

ap = argparse.ArgumentParser()
ap.add_argument('--mode', default='f16x3')
ap.add_argument('--batch', type=int, default=960)
ap.add_argument('--steps', type=int, default=5)
ap.add_argument('--by-shape', action='store_true', help='list GEMMs per shape')
args = ap.parse_args()
net, v, mask_var, window = bench.build_network()
eng = E.get_engine(net, gemm_mode=args.mode)
targets = T.imatrix('targets')
pred = L.get_output(net, deterministic=False)
cost = temporal_softmax_loss(pred, targets, mask_var)
params = L.get_all_params(net, trainable=True)
train = function([v[0], v[1], v[2], targets, mask_var, window], cost, updates=adam(cost, params, learning_rate=1e-3))
xs, mask, y = bench.synth_batch(args.batch, 1000)
dx = [torch.from_numpy(x).cuda() for x in xs]
dmask, dy = torch.from_numpy(mask).cuda(), torch.from_numpy(y).cuda()
for _ in range(3):
    train(dx[0], dx[1], dx[2], dy, dmask, bench.THETA)
prof = E._Profiler(by_shape=args.by_shape)
orig = _lib.call
_lib.call = prof.call
E._lib.call = prof.call
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(args.steps):
    train(dx[0], dx[1], dx[2], dy, dmask, bench.THETA)
e1.record()
summ = prof.summary()
total = e0.elapsed_time(e1) / args.steps
_lib.call = orig
rows = sorted(summ.items(), key=lambda kv: -kv[1][1])
print('mode %s batch %d: %.2f ms/step (with event overhead)' % (args.mode, args.batch, total))
acc = 0.0
for name, (n, t) in rows:
    print('  %-58s n/step=%6.1f  %8.3f ms/step  %5.1f%%' % (name, n / args.steps, t / args.steps, 100 * t / args.steps / total))
    acc += t / args.steps
print('  %-34s %26.3f ms/step' % ('(sum of kernels)', acc))
