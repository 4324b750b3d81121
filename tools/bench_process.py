"""Micro-benchmark of the t-process scan kernels (HBM roofline): python tools/bench_process.py [B ...]"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bruno_b200 import _lib  # noqa: E402
from bruno_b200.nn_extra_student import StudentRecurrentLayer  # noqa: E402
from bruno_b200.nn_extra_gauss import GaussianRecurrentLayer  # noqa: E402
from bruno_b200.process import build_table  # noqa: E402


def timeit(fn, iters=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e-3


def sweep_variants(variants, Bs, T, D, peak):
    for kind, layer in (('tp', StudentRecurrentLayer((D,))), ('gp', GaussianRecurrentLayer((D,)))):
        for B in Bs:
            z = torch.randn(B, T, D, device='cuda')
            nu = layer.nu_vbl.detach() if kind == 'tp' else None
            tab = build_table(layer.kind, layer.var_vbl.detach(), nu, layer.corr_vbl.detach(), layer.mu, None, 2.0, 0, T, False)
            lp = torch.empty(B, T, device='cuda')
            lpp = torch.empty(B, T, device='cuda')
            scratch = torch.empty(max(_lib.query('bruno_process_scratch_floats', B, T, D), 1), device='cuda')
            for v in variants:
                _lib.query('bruno_process_set_fwd_variant', v)
                for name, prior, nbytes in (('fwd', None, 4 * B * T * D + 4 * B * T), ('fwd+prior', lpp, 4 * B * T * D + 8 * B * T)):
                    t = timeit(lambda: _lib.call('bruno_process_fwd', layer.kind, z, layer.mu, tab, None, None, lp, prior, scratch, B, T, D))
                    gbs = nbytes / t * 1e-9
                    print('variant %2d %s %-9s B=%6d T=%d D=%d  %8.1f us  %7.1f GB/s  (%.1f%% of measured %.0f GB/s)' % (
                        v, kind, name, B, T, D, t * 1e6, gbs, 100 * gbs / peak, peak), flush=True)
            _lib.query('bruno_process_set_fwd_variant', 0)
            if os.environ.get('BENCH_BWD_VARIANTS'):
                tabg = build_table(layer.kind, layer.var_vbl.detach(), nu, layer.corr_vbl.detach(), layer.mu, None, 2.0, 0, T, True)
                dz = torch.empty_like(z)
                dv = [torch.empty(1, D, device='cuda') for _ in range(3)]
                bscr = torch.empty(max(_lib.query('bruno_process_bwd_scratch_floats', B, T, D), 1), device='cuda')
                dlp = torch.randn(B, T, device='cuda')
                for v in [int(x) for x in os.environ['BENCH_BWD_VARIANTS'].split(',')]:
                    _lib.query('bruno_process_set_bwd_variant', v)
                    for name, prior in (('bwd', None), ('bwd+prior', dlp)):
                        t = timeit(lambda: _lib.call('bruno_process_bwd', layer.kind, z, layer.mu, tabg, dlp, prior, dz, dv[0],
                                                     dv[1] if kind == 'tp' else None, dv[2], bscr, B, T, D), iters=10)
                        gbs = (8 * B * T * D + 4 * B * T) / t * 1e-9
                        print('bwd variant %2d %s %-9s B=%6d T=%d D=%d  %8.1f us  %7.1f GB/s  (%.1f%% of measured %.0f GB/s)' % (
                            v, kind, name, B, T, D, t * 1e6, gbs, 100 * gbs / peak, peak), flush=True)
                _lib.query('bruno_process_set_bwd_variant', 0)
                del dz


def main():
    peaks = json.load(open(os.path.join(os.path.dirname(__file__), '..', 'MEASURED_PEAKS.json'))) if os.path.exists(
        os.path.join(os.path.dirname(__file__), '..', 'MEASURED_PEAKS.json')) else {'hbm_gbs': 6650.0}
    Bs = [int(a) for a in sys.argv[1:]] or [32, 4096, 16384, 65536]
    T, D = 20, 784
    if os.environ.get('BENCH_VARIANTS'):          # forward-scan A/B: BENCH_VARIANTS=9,10,14 python tools/bench_process.py 65536
        sweep_variants([int(v) for v in os.environ['BENCH_VARIANTS'].split(',')], Bs, T, D, peaks['hbm_gbs'])
        return
    for kind, layer in (('tp', StudentRecurrentLayer((D,))), ('gp', GaussianRecurrentLayer((D,)))):
        for B in Bs:
            z = torch.randn(B, T, D, device='cuda')
            k = layer.kind
            nu = layer.nu_vbl.detach() if kind == 'tp' else None
            tab = build_table(k, layer.var_vbl.detach(), nu, layer.corr_vbl.detach(), layer.mu, None, 2.0, 0, T, True)
            lp = torch.empty(B, T, device='cuda')
            lpp = torch.empty(B, T, device='cuda')
            dz = torch.empty_like(z)
            dv = torch.empty(1, D, device='cuda')
            scratch = torch.empty(max(_lib.query('bruno_process_bwd_scratch_floats', B, T, D), 1), device='cuda')
            for name, fn, nbytes in (
                ('fwd', lambda: _lib.call('bruno_process_fwd', k, z, layer.mu, tab, None, None, lp, None, scratch, B, T, D),
                 4 * B * T * D + 4 * B * T),
                ('fwd+prior', lambda: _lib.call('bruno_process_fwd', k, z, layer.mu, tab, None, None, lp, lpp, scratch, B, T, D),
                 4 * B * T * D + 8 * B * T),
                ('bwd', lambda: _lib.call('bruno_process_bwd', k, z, layer.mu, tab, lp, None, dz, dv, dv if kind == 'gp' else torch.empty_like(dv), dv, scratch, B, T, D),
                 8 * B * T * D + 4 * B * T),
            ):
                t = timeit(fn)
                gbs = nbytes / t * 1e-9
                print('%s %-9s B=%6d T=%d D=%d  %8.1f us  %7.1f GB/s  (%.1f%% of measured %.0f GB/s)' % (
                    kind, name, B, T, D, t * 1e6, gbs, 100 * gbs / peaks['hbm_gbs'], peaks['hbm_gbs']), flush=True)


if __name__ == '__main__':
    main()
