#!/usr/bin/env python
"""Host -> device paths of one padded stream (N x T x F float32, variable lengths): the DMA copy of the whole padded
array, the gather kernel reading pinned host memory directly (all rows / valid rows only), alone and next to a GEMM loop
on another stream.  Prints one JSON line."""
import ctypes as C
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
from ipavsr_b200 import _lib                     # noqa: E402
from ipavsr_b200.engine import _PackPlan         # noqa: E402


def main():
    N, T, F = int(os.environ.get('N', 960)), 40, 1200
    rng = np.random.default_rng(0)
    lens = rng.integers(12, T + 1, size=N)
    plan = _PackPlan(lens, T)
    pin = torch.empty(1 << 20, dtype=torch.int32).pin_memory()
    plan.upload(torch.device('cuda'), pin)
    host = torch.empty(N, T, F, dtype=torch.float32).pin_memory()
    host.normal_()
    dev = torch.empty(N * T, F, device='cuda')
    ident = torch.arange(N * T, dtype=torch.int32, device='cuda')
    cs = torch.cuda.Stream()
    res = {'N': N, 'bytes_padded': N * T * F * 4, 'bytes_valid': int(plan.M) * F * 4}

    def timeit(fn, reps=10, stream=None):
        st = stream or torch.cuda.current_stream()
        with torch.cuda.stream(st):
            for _ in range(2):
                fn(st)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(st)
            for _ in range(reps):
                fn(st)
            e1.record(st)
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps

    def dma(st):
        dev.view(N, T, F).copy_(host, non_blocking=True)

    def gather_all(st):
        _lib.call('ipavsr_gather_rows', host.data_ptr(), 4 * F, dev.data_ptr(), 4 * F, 4 * F, ident.data_ptr(), None,
                  N * T, C.c_void_p(st.cuda_stream))

    def gather_valid(st):
        _lib.call('ipavsr_gather_rows', host.data_ptr(), 4 * F, dev.data_ptr(), 4 * F, 4 * F, plan.pack.data_ptr(), None,
                  plan.M + 1, C.c_void_p(st.cuda_stream))

    for name, fn, nbytes in (('dma_padded', dma, res['bytes_padded']), ('gather_all', gather_all, res['bytes_padded']),
                             ('gather_valid', gather_valid, res['bytes_valid'])):
        ms = timeit(fn, stream=cs)
        res[name] = {'ms': ms, 'GBps': nbytes / ms / 1e6}
    # the same next to a busy GEMM stream
    a = torch.randn(8192, 8192, device='cuda', dtype=torch.bfloat16)
    b = torch.randn(8192, 8192, device='cuda', dtype=torch.bfloat16)
    for name, fn, nbytes in (('dma_padded', dma, res['bytes_padded']), ('gather_valid', gather_valid, res['bytes_valid'])):
        for _ in range(40):
            torch.matmul(a, b)
        ms = timeit(fn, stream=cs, reps=5)
        torch.cuda.synchronize()
        res[name + '_busy'] = {'ms': ms, 'GBps': nbytes / ms / 1e6}
    print(json.dumps(res))


if __name__ == '__main__':
    main()
