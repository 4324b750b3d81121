"""Per-mode timing of the forward/dgrad conv kernel on one or more layer shapes (default: the two bn2_omniglot_tp shapes).
40 back-to-back launches of a dependent chain (out of launch i = in of launch i+1), eager and replayed from ONE CUDA graph
(the 14x14 shape is host-launch-bound in eager mode: the graph number is the GPU-side period of a layer inside a step).
`--sustain S` additionally replays the mode-0 graph of the first shape for S seconds and prints the period per window with
the SM clock nvidia-smi reports and the clock the kernel itself saw (clock64 / globaltimer of CTA 0, DEBUG build).
Usage: python tools/bench_conv.py [N,H,W ...] [--sustain S]"""
import os, subprocess, sys, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bruno_b200 import _lib, slab
args = sys.argv[1:]
sustain = 0.0
if '--sustain' in args:
    i = args.index('--sustain')
    sustain = float(args[i + 1])
    del args[i:i + 2]
shapes = [tuple(int(v) for v in a.split(',')) for a in args] or [(640, 28, 28), (640, 14, 14)]
g = torch.Generator().manual_seed(0)
V = torch.randn(1, 3, 3, 64, 64, generator=g) * 0.05
wf = torch.empty(1, _lib.query('bruno_conv_wpack_bytes'), dtype=torch.uint8, device='cuda')
_lib.call('bruno_conv_wn_pack', V.cuda(), torch.ones(1, 64, device='cuda'), wf, None, 1)
bias = torch.zeros(64, device='cuda')
REP = 40
lib = _lib.load()
if os.environ.get('BRUNO_CONV_VARIANT'): lib.bruno_debug_conv_variant(int(os.environ['BRUNO_CONV_VARIANT']))


def smi():
    q = 'clocks.sm,power.draw,clocks_event_reasons.sw_power_cap'
    return subprocess.run(['nvidia-smi', '-i', '0', '--query-gpu=' + q, '--format=csv,noheader,nounits'],
                          capture_output=True, text=True).stdout.strip()


def timed(fn, rep=REP):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / rep * 1e3


first_graph = None
keep = []
for (N, H, W) in shapes:
    sl = slab.alloc_slabs(5, N, H, W, 'cuda')
    keep.append(sl)
    flops = 2.0 * N * H * W * 9 * 64 * 64
    tiles = (N * (H + 1) * (W + 1) + 127) // 128

    def chain(mode, sl=sl, N=N, H=H, W=W):
        for i in range(REP):                       # ping-pong: launch i reads what launch i-1 wrote
            a, b = sl[i & 1], sl[1 - (i & 1)]
            _lib.call('bruno_conv3x3', mode, a, wf[0], bias if mode in (0, 1) else None, 0, sl[2] if mode else None,
                      sl[3] if mode == 3 else None, b, N, H, W)

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
        us_graph = min(timed(graph.replay) for _ in range(3))
        print('%dx%dx%d (%d tiles, %.1f per CTA) mode %d: eager %.1f us  graph %.1f us per launch  %.0f TFLOP/s (graph)' % (
            N, H, W, tiles, tiles / 148.0, mode, us_eager, us_graph, flops / us_graph * 1e-6), flush=True)
        keep.append(graph)
        if first_graph is None:
            first_graph = []
        if len(first_graph) < 4:
            first_graph.append((graph, flops, (N, H, W), chain, mode))

for graph, flops, shape, chain, gmode in (first_graph if sustain > 0 else []):
    print('sustained replay of mode %d %s for %.1f s; idle: %s' % (gmode, shape, sustain, smi()))
    t_end = time.time() + sustain
    w = 0
    while time.time() < t_end:
        us = timed(lambda: [graph.replay() for _ in range(25)], rep=25 * REP)
        tr = torch.zeros(64 * 16, dtype=torch.int64, device='cuda')
        if hasattr(lib, 'bruno_debug_conv_trace'):
            lib.bruno_debug_conv_trace(tr.data_ptr())
            chain(gmode)
            torch.cuda.synchronize()
            lib.bruno_debug_conv_trace(None)
            t = tr.view(64, 16).cpu()
            cyc, ns = int(t[0, 12] - t[0, 8]), int(t[0, 14] - t[0, 13])
            own = '%d cycles in %d ns = %.0f MHz' % (cyc, ns, 1e3 * cyc / max(ns, 1))
        else:
            own = 'n/a'
        print('  window %2d: %.1f us per launch  %.0f TFLOP/s | smi %s | kernel: %s' % (w, us, flops / us * 1e-6, smi(), own), flush=True)
        w += 1
