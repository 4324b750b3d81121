"""Timeline of CTA 0 of the conv layer chain (forward of one coupling, inference ring): python tools/dbg_trace_chain.py N H W"""
import sys, torch
sys.path.insert(0, '/root/repo')
from bruno_b200 import _lib, nn_extra_nvp
lib = _lib.load()
N, H, W = (int(a) for a in sys.argv[1:4]) if len(sys.argv) > 3 else (640, 14, 14)
layer = nn_extra_nvp.CouplingLayerConv('channel1', name='trace')
x = torch.randn(N, H, W, 4, device='cuda'); ld = torch.zeros(N, device='cuda')
lib.bruno_debug_conv_chain(2)
with torch.no_grad():
    for _ in range(3): layer.forward_and_jacobian(x, None, ld)
    tr = torch.zeros(64 * 16, dtype=torch.int64, device='cuda')
    lib.bruno_debug_conv_trace(tr.data_ptr())
    layer.forward_and_jacobian(x, None, ld)
    torch.cuda.synchronize()
    lib.bruno_debug_conv_trace(None)
t = tr.view(64, 16).cpu(); t0 = int(t[0, 7])
for i in range(40):
    if int(t[i, 0]) == 0: break
    print(i, 'stage free %d load %d | mma start %d issued %d | epi start %d aux ok %d done %d | store %d' % tuple(int(t[i, k] - t0) for k in (7, 0, 1, 2, 3, 6, 4, 5)))
