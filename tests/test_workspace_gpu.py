"""GPU: the boundary contract of include/bruno_sm100.h -- the library never allocates; the dense-path GEMMs take their
split-K workspace from a caller-owned buffer bound per stream, so two streams can run bruno_sgemm concurrently."""
import ctypes

import pytest
import torch

pytestmark = pytest.mark.gpu


def _gemm(lib, st, A, B, C):
    M, K = A.shape
    N = B.shape[1]
    rc = lib.bruno_sgemm(0, 0, M, N, K, A.data_ptr(), K, B.data_ptr(), N, C.data_ptr(), N, None, None, 0, 0, st.cuda_stream)
    assert rc == 0, lib.bruno_last_error()


def test_two_streams_run_split_k_gemms_concurrently():
    from bruno_b200 import _lib
    lib = _lib.load()
    nbytes = lib.bruno_workspace_bytes()
    assert nbytes >= 16 << 20
    streams = [torch.cuda.Stream() for _ in range(2)]
    wss = [torch.zeros(nbytes, dtype=torch.uint8, device='cuda') for _ in range(2)]
    for st, ws in zip(streams, wss):
        assert lib.bruno_workspace_bind(st.cuda_stream, ws.data_ptr(), ws.numel()) == 0
    torch.manual_seed(0)
    # skinny shapes (few output tiles, long K): the split-K path with its tickets and partial tiles
    shapes = [(40, 512, 784), (64, 256, 3072)]
    data = []
    for (M, N, K) in shapes:
        A, B = torch.randn(M, K, device='cuda'), torch.randn(K, N, device='cuda')
        data.append((A, B, torch.empty(M, N, device='cuda'), (A.double() @ B.double())))
    torch.cuda.synchronize()
    for rep in range(50):                                # interleave launches on the two streams
        for st, (A, B, C, ref) in zip(streams, data):
            _gemm(lib, st, A, B, C)
    torch.cuda.synchronize()
    for A, B, C, ref in data:
        assert (C.double() - ref).abs().max().item() <= 2e-5 * ref.abs().max().item()
    for st in streams:
        assert lib.bruno_workspace_unbind(st.cuda_stream) == 0


def test_gemm_without_a_workspace_still_correct():
    """A stream with no bound workspace: the GEMM still splits K (its thread-block-cluster reduction needs no workspace;
    the fallback kernels would run un-split) and nothing is allocated behind the caller's back."""
    from bruno_b200 import _lib
    lib = _lib.load()
    st = torch.cuda.Stream()
    A, B = torch.randn(40, 784, device='cuda'), torch.randn(784, 512, device='cuda')
    C = torch.empty(40, 512, device='cuda')
    _gemm(lib, st, A, B, C)                              # first launch of this kernel variant: the CUDA runtime loads its module
    torch.cuda.synchronize()                             # lazily, which takes device memory once and is not the library's doing
    free0 = torch.cuda.mem_get_info()[0]
    C.zero_()
    _gemm(lib, st, A, B, C)
    torch.cuda.synchronize()
    assert torch.cuda.mem_get_info()[0] == free0         # no device allocation behind the caller's back
    ref = A.double() @ B.double()
    assert (C.double() - ref).abs().max().item() <= 2e-5 * ref.abs().max().item()


def test_bind_rejects_small_or_null_workspace():
    from bruno_b200 import _lib
    lib = _lib.load()
    st = torch.cuda.Stream()
    small = torch.zeros(1024, dtype=torch.uint8, device='cuda')
    assert lib.bruno_workspace_bind(st.cuda_stream, small.data_ptr(), small.numel()) != 0
    assert lib.bruno_workspace_bind(st.cuda_stream, None, lib.bruno_workspace_bytes()) != 0
