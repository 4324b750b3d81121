import sys, torch
sys.path.insert(0, '/root/repo')
from bruno_b200 import _lib
lib = _lib.load()
M, N, K = 640, 512, 784
A = torch.randn(M, K, device='cuda'); B = torch.randn(K, N, device='cuda'); C = torch.empty(M, N, device='cuda')
def run(): _lib.call('bruno_sgemm', 0, 0, M, N, K, A, K, B, N, C, N, None, None, 0, 0)
for _ in range(3): run()
tr = torch.zeros(64 * 8, dtype=torch.int64, device='cuda')
lib.bruno_debug_gemm_trace(tr.data_ptr()); run(); torch.cuda.synchronize(); lib.bruno_debug_gemm_trace(None)
t = tr.view(64, 8).cpu(); t0 = int(t[0, 0])
for i in range(26):
    if int(t[i, 0]) == 0: break
    print(i, 'free %d stored %d arrived %d | issuer saw %d issued %d' % tuple(int(t[i, k]) - t0 for k in range(5)))
print('acc done %d, epilogue end %d' % (int(t[63, 5]) - t0, int(t[63, 6]) - t0))
print('tile in smem %d, rows written %d' % (int(t[60, 5]) - t0, int(t[60, 6]) - t0))
