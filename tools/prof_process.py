import sys, torch
sys.path.insert(0, '/root/repo')
from bruno_b200 import _lib
from bruno_b200.nn_extra_student import StudentRecurrentLayer
from bruno_b200.process import build_table
B, T, D = 32768, 20, 784
layer = StudentRecurrentLayer((D,))
z = torch.randn(B, T, D, device='cuda')
tab = build_table(0, layer.var_vbl.detach(), layer.nu_vbl.detach(), layer.corr_vbl.detach(), layer.mu, None, 2.0, 0, T, False)
lp = torch.empty(B, T, device='cuda')
lpp = torch.empty(B, T, device='cuda') if 'prior' in sys.argv[1:] else None   # python tools/prof_process.py prior
sc = torch.empty(1, device='cuda')
for _ in range(3):
    _lib.call('bruno_process_fwd', 0, z, layer.mu, tab, None, None, lp, lpp, sc, B, T, D)
torch.cuda.synchronize()
print('ok')
