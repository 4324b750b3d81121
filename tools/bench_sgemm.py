import sys, torch
sys.path.insert(0, '/root/repo')
from bruno_b200 import _lib
def t(fn, it=50):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(20): fn()
    run = fn
    fn = g.replay
    it = 10
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(it): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / it * 1e3 / 20
for (ta, tb, M, N, K) in [(0, 0, 640, 512, 784), (0, 0, 640, 784, 512), (1, 0, 784, 512, 640), (0, 1, 640, 512, 784), (0, 0, 1280, 512, 784), (0, 0, 40, 512, 784)]:
    A = torch.randn((K, M) if ta else (M, K), device='cuda')
    B = torch.randn((N, K) if tb else (K, N), device='cuda')
    C = torch.empty(M, N, device='cuda')
    us = t(lambda: _lib.call('bruno_sgemm', ta, tb, M, N, K, A, A.shape[1], B, B.shape[1], C, N, None, None, 0, 0))
    ref = t(lambda: torch.matmul(A.t() if ta else A, B.t() if tb else B))
    print('ta=%d tb=%d M=%d N=%d K=%d: %.1f us = %.1f TFLOP/s   (torch.matmul fp32: %.1f us)' % (ta, tb, M, N, K, us, 2 * M * N * K / us * 1e-6, ref))
