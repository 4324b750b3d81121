"""Per-mode timing of the forward/dgrad conv kernel on one layer shape (default: the bn2_omniglot_tp 28x28 layer).
40 back-to-back launches of a dependent chain (out of launch i = in of launch i+1), eager and replayed from ONE CUDA graph
(the 14x14 shape is host-launch-bound in eager mode: the graph number is the GPU-side period of a layer inside a step).
Usage: python tools/bench_conv.py [N H W]"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bruno_b200 import _lib, slab
N, H, W = 640, 28, 28
if len(sys.argv) > 3: N, H, W = (int(a) for a in sys.argv[1:4])
g = torch.Generator().manual_seed(0)
V = torch.randn(1, 3, 3, 64, 64, generator=g) * 0.05
wf = torch.empty(1, _lib.query('bruno_conv_wpack_bytes'), dtype=torch.uint8, device='cuda')
_lib.call('bruno_conv_wn_pack', V.cuda(), torch.ones(1, 64, device='cuda'), wf, None, 1)
sl = slab.alloc_slabs(5, N, H, W, 'cuda')
bias = torch.zeros(64, device='cuda')
flops = 2.0 * N * H * W * 9 * 64 * 64
REP = 40


def chain(mode):
    for i in range(REP):                       # ping-pong: launch i reads what launch i-1 wrote
        a, b = sl[i & 1], sl[1 - (i & 1)]
        _lib.call('bruno_conv3x3', mode, a, wf[0], bias if mode in (0, 1) else None, 0, sl[2] if mode else None,
                  sl[3] if mode == 3 else None, b, N, H, W)


def timed(fn):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / REP * 1e3


for mode in (0, 1, 2, 3):
    chain(mode)
    us_eager = timed(lambda: chain(mode))
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        chain(mode)
    torch.cuda.current_stream().wait_stream(s)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph, stream=s):
        chain(mode)
    graph.replay()
    us_graph = timed(graph.replay)
    print('mode %d: eager %.1f us  graph %.1f us per launch  %.0f TFLOP/s (graph)' % (mode, us_eager, us_graph, flops / us_graph * 1e-6))
