"""GPU-side period of the dense-path GEMM (bruno_sgemm, 3xTF32 tcgen05 kernel) at the shapes of one dense coupling of
bn2_omniglot_tp (640 rows = 32 sequences x 20 frames, D = 784, U = 512): 20 launches per shape replayed from ONE CUDA graph
captured on a stream whose split-K workspace is bound (an unbound stream runs un-split: a different kernel shape), next to
torch.matmul (cuBLAS fp32) captured the same way.  BRUNO_TF32_SPLIT / BRUNO_TF32_BN force the split / tile width.
Usage: python tools/bench_sgemm.py"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bruno_b200 import _lib

REP = 20
side = torch.cuda.Stream()


def graph_us(fn):
    with torch.cuda.stream(side):                       # warm-up on the capture stream binds its workspace
        for _ in range(3):
            fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=side):
        for _ in range(REP):
            fn()
    g.replay()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            g.replay()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / (10 * REP) * 1e3)
    return best


# (ta, tb, M, N, K, what)
SHAPES = [(0, 0, 640, 512, 784, 'fwd d1'), (0, 0, 640, 512, 512, 'fwd d2'), (0, 0, 640, 784, 512, 'fwd head'),
          (1, 0, 512, 784, 640, 'bwd dK head'), (0, 1, 640, 512, 784, 'bwd dh from head'), (1, 0, 512, 512, 640, 'bwd dV2'),
          (0, 1, 640, 512, 512, 'bwd dh1'), (1, 0, 784, 512, 640, 'bwd dV1'), (0, 1, 640, 784, 512, 'bwd dxm'),
          (0, 0, 40, 512, 784, 'few-shot d1')]
total = 0.0
for (ta, tb, M, N, K, what) in SHAPES:
    A = torch.randn((K, M) if ta else (M, K), device='cuda')
    B = torch.randn((N, K) if tb else (K, N), device='cuda')
    C = torch.empty(M, N, device='cuda')
    us = graph_us(lambda: _lib.call('bruno_sgemm', ta, tb, M, N, K, A, A.shape[1], B, B.shape[1], C, N, None, None, 0, 0))
    ref = graph_us(lambda: torch.matmul(A.t() if ta else A, B.t() if tb else B, out=C))
    total += us
    print('%-18s ta=%d tb=%d M=%4d N=%4d K=%4d: %5.1f us = %5.1f TFLOP/s   (torch.matmul fp32: %5.1f us)' % (
        what, ta, tb, M, N, K, us, 2 * M * N * K / us * 1e-6, ref), flush=True)
print('sum over the listed shapes: %.1f us' % total)
