"""Host-side view of the activation-slab layout (csrc/slab.cuh): geometry, allocation, and NHWC <-> slab
conversion helpers.  The conversions are for tests / debugging / data-dependent init only -- the hot path never
leaves the slab layout."""
import torch

from . import _lib

TILE_M = 128


class SlabGeom(object):
    def __init__(self, N, H, W):
        self.N, self.H, self.W = N, H, W
        p8 = ((W + 1 + 7) // 8) * 8
        self.P = p8 if 4 * p8 <= 5 * (W + 1) else W + 1          # csrc/slab.cuh: slab_pitch
        self.IMG = (H + 1) * self.P
        self.HALO = ((self.P + 1 + 7) // 8) * 8
        self.data_rows = N * self.IMG
        self.NT = (self.data_rows + TILE_M - 1) // TILE_M
        self.rows = (self.NT + 2) * TILE_M + 2 * self.HALO

    @property
    def nbytes(self):
        return self.rows * 128


def alloc_slabs(n, N, H, W, device):
    """n zero-initialised slabs as one [n, bytes] uint8 tensor (the kernels keep padding rows zero)."""
    nbytes = _lib.query('bruno_slab_bytes', N, H, W)
    assert nbytes == SlabGeom(N, H, W).nbytes
    return torch.zeros(n, nbytes, dtype=torch.uint8, device=device)


def _pixel_rows(g, device):
    n = torch.arange(g.N, device=device).view(-1, 1, 1)
    y = torch.arange(g.H, device=device).view(1, -1, 1)
    x = torch.arange(g.W, device=device).view(1, 1, -1)
    return (g.HALO + n * g.IMG + (y + 1) * g.P + (x + 1)).reshape(-1)


def _swizzle(rows_bf16):
    """[rows, 64] bf16 -> same with 16-byte chunk j of row R moved to chunk j ^ (R & 7) (an involution)."""
    R = rows_bf16.shape[0]
    v = rows_bf16.view(R, 8, 8)
    r = torch.arange(R, device=v.device).view(R, 1) & 7
    src = torch.arange(8, device=v.device).view(1, 8) ^ r
    return torch.gather(v, 1, src.view(R, 8, 1).expand(R, 8, 8)).reshape(R, 64)


def nhwc_to_slab(x):
    """x [N,H,W,64] float -> uint8 slab."""
    N, H, W, C = x.shape
    assert C == 64
    g = SlabGeom(N, H, W)
    buf = torch.zeros(g.rows, 64, dtype=torch.bfloat16, device=x.device)
    buf[_pixel_rows(g, x.device)] = x.reshape(-1, 64).to(torch.bfloat16)
    return _swizzle(buf).contiguous().view(torch.uint8).reshape(-1)


def slab_to_nhwc(slab, N, H, W, check_padding=False):
    g = SlabGeom(N, H, W)
    buf = _swizzle(slab.view(torch.bfloat16).reshape(g.rows, 64))
    rows = _pixel_rows(g, slab.device)
    if check_padding:
        m = torch.ones(g.rows, dtype=torch.bool, device=slab.device)
        m[rows] = False
        assert float(buf[m].float().abs().max()) == 0.0, 'slab padding rows are not zero'
    return buf[rows].float().reshape(N, H, W, 64)
