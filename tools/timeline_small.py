"""Kernel timeline of the small-batch (CUDA-graph) training step from torch.profiler (CUPTI): which kernels overlap,
what the critical chain is.   python tools/timeline_small.py [--batch 26]"""
import argparse, os, sys, collections
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from ipavsr_b200 import layers as L
from ipavsr_b200.function import function, tensor as T
from ipavsr_b200.custom.objectives import temporal_softmax_loss
from ipavsr_b200.custom.updates import adam

ap = argparse.ArgumentParser()
ap.add_argument('--batch', type=int, default=26)
args = ap.parse_args()
net, v, mask_var, window = bench.build_network()
targets = T.imatrix('targets')
cost = temporal_softmax_loss(L.get_output(net, deterministic=False), targets, mask_var)
params = L.get_all_params(net, trainable=True)
train = function([v[0], v[1], v[2], targets, mask_var, window], cost, updates=adam(cost, params, learning_rate=1e-3))
xs, mask, y = bench.synth_batch(args.batch, 1000)
dx = [torch.from_numpy(x).cuda() for x in xs]
dmask, dy = torch.from_numpy(mask).cuda(), torch.from_numpy(y).cuda()
for _ in range(5):
    train(dx[0], dx[1], dx[2], dy, dmask, bench.THETA)
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(3):
        train(dx[0], dx[1], dx[2], dy, dmask, bench.THETA)
    torch.cuda.synchronize()
import json, tempfile
tmp = os.path.join(tempfile.gettempdir(), 'ipavsr_trace.json')
prof.export_chrome_trace(tmp)
tr = json.load(open(tmp))
ev = [e for e in tr['traceEvents'] if e.get('cat') == 'kernel']
ev.sort(key=lambda e: e['ts'])
opt = [i for i, e in enumerate(ev) if 'optim_kernel' in e['name']]
lo = opt[-2] + 1 if len(opt) >= 2 else 0
step = ev[lo:opt[-1] + 1]
t0 = step[0]['ts']
end = max(e['ts'] + e['dur'] for e in step)
print('%d kernels, span %.1f us, sum of durations %.1f us' % (len(step), end - t0, sum(e['dur'] for e in step)))
pts = []
for e in step:
    pts.append((e['ts'], 1)); pts.append((e['ts'] + e['dur'], -1))
pts.sort()
cur, last, hist = 0, t0, collections.Counter()
for t, d in pts:
    hist[cur] += t - last; last = t; cur += d
print('time with k kernels running (us):', {k: round(v, 1) for k, v in sorted(hist.items())})
streams = sorted({e['args'].get('stream') for e in step})
print('streams:', streams)
short = lambda n: n.replace('void ', '').replace('ipavsr::', '').replace('(anonymous namespace)::', '').split('(')[0][:44]
prev_end = t0
for e in step:
    gap = e['ts'] - prev_end
    print('%9.1f %8.1f  s%-3d %s%s' % (e['ts'] - t0, e['dur'], streams.index(e['args'].get('stream')), short(e['name']),
                                      ('   <- idle %.1f us before' % gap) if gap > 8 else ''))
    prev_end = max(prev_end, e['ts'] + e['dur'])
