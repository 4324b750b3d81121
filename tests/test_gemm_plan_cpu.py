"""Host logic of the dense-path GEMM launcher (no GPU): tile width, split-K slices = thread-block cluster size, paired products.
bruno_debug_gemm_plan runs the same functions the launcher runs (csrc/gemm_tf32.cu: plan_bn / plan_split)."""
import ctypes
import random

import pytest


def plan(M, N, K, mode=0, sms=148):
    from bruno_b200 import _lib
    lib = _lib.load()
    if not hasattr(lib, 'bruno_debug_gemm_plan'):
        pytest.skip('library built with -DBRUNO_NO_DEBUG')
    out = (ctypes.c_int * 6)()
    assert lib.bruno_debug_gemm_plan(M, N, K, mode, sms, out) == 0
    return dict(zip(('bn', 'gx', 'gy', 'gz', 'kchunk', 'cluster'), list(out)))


def test_shapes_of_one_dense_coupling_on_148_sms():
    # bn2_omniglot_tp: 640 rows (32 sequences x 20 frames), D = 784, U = 512
    assert plan(640, 512, 784) == dict(bn=64, gx=8, gy=5, gz=3, kchunk=9, cluster=3)                 # d1: 120 CTAs
    assert plan(640, 512, 512) == dict(bn=64, gx=8, gy=5, gz=3, kchunk=6, cluster=3)                 # d2
    assert plan(40, 512, 784) == dict(bn=64, gx=8, gy=1, gz=7, kchunk=4, cluster=7)                  # few-shot episode
    # s | t heads side by side: 130 tiles, no split-K at all; dKs | dKt: 104 tiles
    assert plan(640, 784, 512, mode=1) == dict(bn=64, gx=26, gy=5, gz=1, kchunk=16, cluster=1)
    assert plan(512, 784, 640, mode=1) == dict(bn=64, gx=26, gy=4, gz=1, kchunk=20, cluster=1)
    # dh = ds Ks^T + dt Kt^T: one slice per product, a cluster of two
    assert plan(640, 512, 784, mode=2) == dict(bn=64, gx=8, gy=5, gz=2, kchunk=25, cluster=2)
    # a large product: 128-column tiles, un-split, no cluster; pairs then need two launches (gz = 0)
    big = plan(8192, 8192, 512)
    assert (big['bn'], big['gz'], big['cluster']) == (128, 1, 1)
    assert plan(8192, 8192, 512, mode=2)['gz'] == 0


def test_plan_invariants_on_random_shapes():
    rng = random.Random(7)
    for _ in range(2000):
        M, N, K = rng.randint(1, 3000), 4 * rng.randint(1, 600), 4 * rng.randint(1, 600)
        mode, sms = rng.choice((0, 0, 1, 2)), rng.choice((148, 132, 80, 16))
        p = plan(M, N, K, mode, sms)
        if p['gz'] == 0:                                   # pair on 128-column tiles: the caller launches the products one by one
            assert mode and p['bn'] == 128
            continue
        nk = -(-K // 32)
        per_product = p['gz'] // 2 if mode == 2 else p['gz']
        assert p['bn'] in (64, 128) and p['gy'] == -(-M // 128) and p['gx'] == -(-N // p['bn']) * (2 if mode == 1 else 1)
        assert 1 <= p['gz'] <= 8 and (mode != 2 or p['gz'] % 2 == 0)
        assert per_product * p['kchunk'] >= nk             # the slices cover K ...
        assert (per_product - 1) * p['kchunk'] < nk        # ... and none is empty
        assert per_product == 1 or p['kchunk'] >= 2        # at least two K tiles per CTA when K is split
        assert p['cluster'] == (p['gz'] if p['bn'] == 64 and p['gz'] > 1 else 1)
        if p['cluster'] > 1:                               # the push buffer holds S x ceil(rows / S) partial rows
            assert p['cluster'] * -(-128 // p['cluster']) <= 128 + 8
