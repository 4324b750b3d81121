"""Timeline of CTA (0,0,0) of the 3xTF32 GEMM (bruno_debug_gemm_trace): per K tile: stage free, tile stored, arrived, issuer
saw it, MMAs issued; row 63: accumulator done, epilogue end.  Usage: python tools/trace_gemm.py [ta tb M N K]"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bruno_b200 import _lib
ta, tb, M, N, K = (int(a) for a in sys.argv[1:6]) if len(sys.argv) > 5 else (0, 0, 640, 512, 784)
A = torch.randn((K, M) if ta else (M, K), device='cuda'); B = torch.randn((N, K) if tb else (K, N), device='cuda')
C = torch.empty(M, N, device='cuda')
lib = _lib.load()
f = lambda: _lib.call('bruno_sgemm', ta, tb, M, N, K, A, A.shape[1], B, B.shape[1], C, N, None, None, 0, 0)
for _ in range(3): f()
tr = torch.zeros(64 * 8, dtype=torch.int64, device='cuda')
lib.bruno_debug_gemm_trace(tr.data_ptr()); f(); torch.cuda.synchronize(); lib.bruno_debug_gemm_trace(None)
t = tr.view(64, 8).cpu()
base = min(int(v) for v in t.flatten() if int(v) > 0)
for i in range(64):
    if t[i].abs().sum() == 0: continue
    print('row %2d: ' % i + ' '.join('%7d' % (int(v) - base if int(v) > 0 else -1) for v in t[i]))
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(50): f()
e1.record(); torch.cuda.synchronize()
print('%.1f us per GEMM (eager back-to-back)' % (e0.elapsed_time(e1) / 50 * 1e3))
