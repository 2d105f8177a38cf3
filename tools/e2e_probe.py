import sys, time, os
sys.path.insert(0, '/root/repo')
import numpy as np, torch
import bench
from ipavsr_b200 import layers as L
from ipavsr_b200.function import function, tensor as T
from ipavsr_b200.custom.objectives import temporal_softmax_loss
from ipavsr_b200.custom.updates import adam
from ipavsr_b200.derived import DiffImages, DctFeatures
net, v, mask_var, window = bench.build_network()
targets = T.imatrix('targets')
cost = temporal_softmax_loss(L.get_output(net, deterministic=False), targets, mask_var)
params = L.get_all_params(net, trainable=True)
train = function([v[0], v[1], v[2], targets, mask_var, window], cost, updates=adam(cost, params, learning_rate=1e-3))
B = 512
pin = lambda a: torch.from_numpy(a).pin_memory()
host, dev = [], []
for b in range(4):
    xs, mask, y = bench.synth_batch(B, 1000 + b)
    host.append((pin(xs[0]), pin(mask), pin(y)))
    dev.append(([torch.from_numpy(x).cuda() for x in xs], torch.from_numpy(mask).cuda(), torch.from_numpy(y).cuda()))
def timeit(fn, n=20, w=4):
    for i in range(w): fn(i)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for i in range(n): fn(w + i)
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / n * 1e3
def dev_step(i):
    xs, m, y = dev[i % 4]; train(xs[0], xs[1], xs[2], y, m, 9)
def dev_derived(i):
    xs, m, y = dev[i % 4]; train(xs[0], DiffImages(xs[0]), DctFeatures(xs[0], (30, 40), 30), y, m, 9)
args = [(h[0], DiffImages(h[0]), DctFeatures(h[0], (30, 40), 30), h[2], h[1], 9) for h in host]
def e2e_nopf(i):
    train(*args[i % 4])
def e2e_pf(i):
    train.prefetch(*args[(i + 1) % 4]); train(*args[i % 4])
def e2e_defer(i):
    train.prefetch(*args[(i + 1) % 4], defer=True); train(*args[i % 4])
print('device 3 streams        %.3f ms' % timeit(dev_step))
print('device raw + derived    %.3f ms' % timeit(dev_derived))
print('host raw, no prefetch   %.3f ms' % timeit(e2e_nopf))
train.prefetch(*args[0])
print('host raw, prefetch      %.3f ms' % timeit(e2e_pf))
train.engine._prefetched = []
train.prefetch(*args[0])
print('host raw, deferred prefetch %.3f ms' % timeit(e2e_defer))
train.engine._prefetched = []
t0 = time.perf_counter()
for i in range(20): train.prefetch(*args[i % 4]); train.engine._prefetched = []
torch.cuda.synchronize()
print('prefetch call alone (host+gather)  %.3f ms' % ((time.perf_counter() - t0) / 20 * 1e3))
