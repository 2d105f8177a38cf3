"""K training steps of the bench workload (device-resident inputs) for ncu launch lists:
    ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches.csv \
        python tools/ncu_step.py [--batch 512] [--steps 3]
The step boundaries are marked by a tiny `ipavsr_fill` of 1 element with the value 12345 (searchable in the list)."""
import argparse, ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from ipavsr_b200 import layers as L, _lib, engine as E
from ipavsr_b200.function import function, tensor as T
from ipavsr_b200.custom.objectives import temporal_softmax_loss
from ipavsr_b200.custom.updates import adam

ap = argparse.ArgumentParser()
ap.add_argument('--batch', type=int, default=512)
ap.add_argument('--steps', type=int, default=3)
args = ap.parse_args()
net, v, mask_var, window = bench.build_network()
targets = T.imatrix('targets')
cost = temporal_softmax_loss(L.get_output(net, deterministic=False), targets, mask_var)
params = L.get_all_params(net, trainable=True)
train = function([v[0], v[1], v[2], targets, mask_var, window], cost, updates=adam(cost, params, learning_rate=1e-3))
xs, mask, y = bench.synth_batch(args.batch, 1000)
dx = [torch.from_numpy(x).cuda() for x in xs]
dmask, dy = torch.from_numpy(mask).cuda(), torch.from_numpy(y).cuda()
marker = torch.zeros(4, device='cuda')
for i in range(args.steps):
    _lib.call('ipavsr_fill', marker.data_ptr(), 1, 12345.0, C.c_void_p(torch.cuda.current_stream().cuda_stream))
    train(dx[0], dx[1], dx[2], dy, dmask, bench.THETA)
torch.cuda.synchronize()
print('done')
