"""Host time per training step: from the call of the compiled function to the point where it blocks on the loss
(everything the host has to do to keep the device fed), device-resident inputs vs the end-to-end path (pinned host raw
stream, derived streams, deferred prefetch), with per-call wall times and the allocator's cudaMalloc count.
    python tools/host_time.py [batch]        IPAVSR_EARLY_LOSS=0: late loss read;  HOST_TIME_DEFER=0: prefetch(..., defer=False)"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import bench
from ipavsr_b200 import layers as L
from ipavsr_b200.derived import DiffImages, DctFeatures
from ipavsr_b200.function import function, tensor as T
from ipavsr_b200.custom.objectives import temporal_softmax_loss
from ipavsr_b200.custom.updates import adam

B = int(sys.argv[1]) if len(sys.argv) > 1 else 512
net, v, mask_var, window = bench.build_network()
targets = T.imatrix('targets')
cost = temporal_softmax_loss(L.get_output(net, deterministic=False), targets, mask_var)
params = L.get_all_params(net, trainable=True)
train = function([v[0], v[1], v[2], targets, mask_var, window], cost, updates=adam(cost, params, learning_rate=1e-3))
eng = train.engine
marks = []
orig = eng.read_loss
def read_loss():
    marks.append(time.perf_counter())
    return orig()
eng.read_loss = read_loss
pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
NB = 4
dev, host = [], []
for b in range(NB):
    xs, mask, y = bench.synth_batch(B, 100 + b)
    dev.append(([torch.from_numpy(x).cuda() for x in xs], torch.from_numpy(mask).cuda(), torch.from_numpy(y).cuda()))
    host.append((pin(xs[0]), pin(mask), pin(y)))

def run(step, n=24, warm=4):
    host_ms, wall = [], []
    torch.cuda.synchronize()
    for i in range(n + warm):
        t0 = time.perf_counter()
        step(i)
        t1 = time.perf_counter()
        if i >= warm:
            host_ms.append((marks[-1] - t0) * 1e3)
            wall.append((t1 - t0) * 1e3)
    torch.cuda.synchronize()
    print('   per-call wall (ms):', ' '.join('%.1f' % w for w in wall))
    return np.median(host_ms), np.median(wall)

h, w = run(lambda i: train(dev[i % NB][0][0], dev[i % NB][0][1], dev[i % NB][0][2], dev[i % NB][2], dev[i % NB][1], bench.THETA))
print('batch %d device-resident: host %.2f ms of a %.2f ms step' % (B, h, w))
args = [(r, DiffImages(r), DctFeatures(r, bench.IMAGE_SHAPE, bench.DCT_COEFF), y, m, bench.THETA) for (r, m, y) in host]
train.prefetch(*args[0])
DEFER = os.environ.get('HOST_TIME_DEFER', '1') == '1'     # 0: the next batch is staged BEFORE this call enqueues its kernels
def e2e(i):
    train.prefetch(*args[(i + 1) % NB], defer=DEFER)
    train(*args[i % NB])
ms0 = torch.cuda.memory_stats()
h, w = run(e2e)
ms1 = torch.cuda.memory_stats()
print('   cudaMalloc calls during the loop: %d, cudaFree: %d, reserved %.0f -> %.0f MB' % (
    ms1.get('num_device_alloc', 0) - ms0.get('num_device_alloc', 0), ms1.get('num_device_free', 0) - ms0.get('num_device_free', 0),
    ms0['reserved_bytes.all.current'] / 2**20, ms1['reserved_bytes.all.current'] / 2**20))
print('batch %d end to end:      host %.2f ms of a %.2f ms step' % (B, h, w))
if os.environ.get('HOST_TIME_PROFILE') == '1':
    import cProfile, pstats
    pr = cProfile.Profile(); pr.enable()
    for i in range(12): e2e(i)
    pr.disable()
    pstats.Stats(pr).sort_stats('cumulative').print_stats(32)
