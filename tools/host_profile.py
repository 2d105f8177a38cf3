"""cProfile of the host side of a training step of the bench network (python tools/host_profile.py [batch])."""
import sys, os, time, cProfile, pstats
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
import torch
import bench
from ipavsr_b200 import layers as L
from ipavsr_b200.function import function, tensor as T
from ipavsr_b200.custom.objectives import temporal_softmax_loss
from ipavsr_b200.custom.updates import adam
B = int(sys.argv[1]) if len(sys.argv) > 1 else 26
net, v, mask_var, window = bench.build_network()
targets = T.imatrix('targets')
cost = temporal_softmax_loss(L.get_output(net, deterministic=False), targets, mask_var)
params = L.get_all_params(net, trainable=True)
train = function([v[0], v[1], v[2], targets, mask_var, window], cost, updates=adam(cost, params, learning_rate=1e-3))
xs, mask, y = bench.synth_batch(B, 1)
dx = [torch.from_numpy(x).cuda() for x in xs]
dm, dy = torch.from_numpy(mask).cuda(), torch.from_numpy(y).cuda()
for _ in range(5):
    train(dx[0], dx[1], dx[2], dy, dm, 9)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(20):
    train(dx[0], dx[1], dx[2], dy, dm, 9)
torch.cuda.synchronize()
print('batch %d: %.3f ms/step wall' % (B, (time.perf_counter() - t0) / 20 * 1e3))
pr = cProfile.Profile()
pr.enable()
for _ in range(20):
    train(dx[0], dx[1], dx[2], dy, dm, 9)
pr.disable()
st = pstats.Stats(pr)
st.sort_stats('tottime').print_stats(28)
