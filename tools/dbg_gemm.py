"""Accuracy of the dense-path GEMM variants against float64 for all transpose combinations and edge shapes."""
import sys, torch
sys.path.insert(0, '/root/repo')
from bruno_b200 import _lib
lib = _lib.load()
torch.manual_seed(0)
def run(ta, tb, M, N, K, variant, epi=False):
    lib.bruno_debug_gemm_variant(variant)
    A = torch.randn((K, M) if ta else (M, K), device='cuda')
    B = torch.randn((N, K) if tb else (K, N), device='cuda')
    C = torch.full((M, N), 7.0, device='cuda')
    cs = torch.rand(N, device='cuda') + 0.5 if epi else None
    bias = torch.randn(N, device='cuda') if epi else None
    _lib.call('bruno_sgemm', ta, tb, M, N, K, A, A.shape[1], B, B.shape[1], C, N, cs, bias, 1 if epi else 0, 0)
    torch.cuda.synchronize()
    ref = (A.double().t() if ta else A.double()) @ (B.double().t() if tb else B.double())
    if epi:
        ref = torch.nn.functional.elu(ref * cs.double() + bias.double())
    err = (C.double() - ref).abs().max().item() / ref.abs().max().item()
    return err
shapes = [(640, 512, 784), (640, 784, 512), (784, 512, 640), (40, 512, 784), (128, 128, 32), (132, 260, 100), (1, 4, 4), (300, 64, 64)]
for (M, N, K) in shapes:
    for ta in (0, 1):
        for tb in (0, 1):
            if (ta and M % 4) or ((not tb) and N % 4) or K % 4: continue
            e1 = run(ta, tb, M, N, K, 1); e0 = run(ta, tb, M, N, K, 0); e1e = run(ta, tb, M, N, K, 1, True)
            flag = 'OK ' if e1 < 5e-6 and e1e < 5e-6 else 'BAD'
            print('%s M=%d N=%d K=%d ta=%d tb=%d: tf32x3 %.2e  (epilogue %.2e)   simt %.2e' % (flag, M, N, K, ta, tb, e1, e1e, e0))
