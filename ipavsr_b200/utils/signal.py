"""`utils/signal.py:59-80` (`append_delta_coeff`) as a host-callable function backed by the DeltaLayer kernel."""
import ctypes as C

import numpy as np
import torch

from .. import _lib


def append_delta_coeff(A, theta, exact=True):
    """A: (T, F) or (N, T, F) float32 -> [A | delta | accel] with window half-width theta (edge replicated)."""
    a = np.asarray(A, dtype=np.float32)
    single = a.ndim == 2
    if single:
        a = a[None]
    N, T, F = a.shape
    ldx, ldy = (F + 3) // 4 * 4, (3 * F + 3) // 4 * 4
    xp = np.zeros((N * T, ldx), np.float32)
    xp[:, :F] = a.reshape(N * T, F)
    x = torch.from_numpy(xp).cuda()
    y = torch.zeros(N * T, ldy, dtype=torch.float32, device='cuda')
    _lib.call('ipavsr_delta_fwd', x.data_ptr(), ldx, y.data_ptr(), ldy, N, T, F, int(theta), 1 if exact else 0,
              C.c_void_p(torch.cuda.current_stream().cuda_stream))
    out = y.cpu().numpy()[:, :3 * F].reshape(N, T, 3 * F)
    return out[0] if single else out
