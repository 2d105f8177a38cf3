"""Phase breakdown of the tensor-core LSTM forward (cycles accumulated by CTA 0)."""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from ipavsr_b200 import _lib
lib = _lib.load()
st = lambda: C.c_void_p(torch.cuda.current_stream().cuda_stream)
N, T, H = int(sys.argv[1]) if len(sys.argv) > 1 else 480, 40, 250
ldh = 256
xw = torch.randn(N * T, 4 * H, device='cuda'); whid = torch.randn(H, 4 * H, device='cuda') * 0.05
peep = torch.randn(3, H, device='cuda') * 0.1; z = torch.zeros(H, device='cuda')
lens = torch.randint(12, T + 1, (N,), device='cuda')
mask = (torch.arange(T, device='cuda')[None, :] < lens[:, None]).to(torch.uint8).contiguous()
out, hprev = torch.zeros(N * T, ldh, device='cuda'), torch.zeros(N * T, ldh, device='cuda')
gates, cell = torch.empty(N * T, 4 * H, device='cuda'), torch.empty(N * T, H, device='cuda')
wh, wl = torch.empty(H, 4 * H, dtype=torch.float16, device='cuda'), torch.empty(H, 4 * H, dtype=torch.float16, device='cuda')
sc = torch.zeros(2, device='cuda')
_lib.call('ipavsr_f16_split', whid.data_ptr(), 4 * H, H, 4 * H, wh.data_ptr(), wl.data_ptr(), 4 * H, sc.data_ptr(), sc.data_ptr() + 4, 0, st())
run = lambda: _lib.call('ipavsr_lstm_fwd_f16', xw.data_ptr(), wh.data_ptr(), wl.data_ptr(), sc.data_ptr() + 4, 4 * H, peep.data_ptr(), z.data_ptr(),
                        z.data_ptr(), mask.data_ptr(), out.data_ptr(), gates.data_ptr(), cell.data_ptr(), hprev.data_ptr(), N, T, H, ldh, 0, st())
for _ in range(3): run()
buf = torch.zeros(8, dtype=torch.int64, device='cuda')
lib.ipavsr_debug_lstm_timestamps(C.c_void_p(buf.data_ptr()))
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); run(); e1.record(); torch.cuda.synchronize()
lib.ipavsr_debug_lstm_timestamps(None)
b = buf.cpu().numpy()
print('N=%d: %.3f ms, %.2f us/step' % (N, e0.elapsed_time(e1), e0.elapsed_time(e1) * 1e3 / T))
names = ['control: wait h', 'control: MMA issue', 'epilogue: wait acc', 'epilogue: TMEM ld', 'epilogue: math+stores', 'epilogue: push h']
for n, v in zip(names, b[:6]):
    print('  %-24s %8.0f cycles/step' % (n, v / T))
