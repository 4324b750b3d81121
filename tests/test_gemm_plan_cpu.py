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


def test_cluster_reduction_ownership_model():
    """Python model of the split-K reduction inside a cluster (csrc/gemm_tf32.cu, epilogue): CTA z owns rows
    [z rows_per, (z + 1) rows_per) of the 128-row tile; every CTA pushes each of its partial rows into slot
    [source z][local row] of the owner's buffer and the owner waits for S x (rows it owns) x 256 bytes.  Every row has exactly
    one owner, the byte counts add up, and the slots of one owner never overlap or leave the 136-row buffer."""
    BN, PUSH_ROWS = 64, 128 + 8
    for S in range(2, 9):
        for rows in list(range(1, 129)):
            rows_per = -(-rows // S)
            expected = [S * max(0, min(rows, (z + 1) * rows_per) - z * rows_per) * BN * 4 for z in range(S)]
            received = [0] * S
            slots = [set() for _ in range(S)]
            for src in range(S):
                for r in range(128):                      # TMEM lane = tile row
                    if r >= rows:
                        continue
                    owner, lr = r // rows_per, r % rows_per
                    assert owner < S
                    slot = src * rows_per + lr
                    assert slot < PUSH_ROWS and slot not in slots[owner]
                    slots[owner].add(slot)
                    received[owner] += BN * 4
            assert received == expected
            assert sum(expected) == S * rows * BN * 4
            # the reduction of owner z reads slots z' * rows_per + lr for z' < S: exactly the slots written
            for z in range(S):
                nmine = max(0, min(rows, (z + 1) * rows_per) - z * rows_per)
                assert slots[z] == {zz * rows_per + lr for zz in range(S) for lr in range(nmine)}
