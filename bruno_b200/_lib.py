"""ctypes binding of libbruno_sm100.so (the C ABI declared in include/bruno_sm100.h).

There is no CPU fallback: if the library is missing or a call fails this module raises.  Prototypes are parsed
from the header so the Python side can never drift from the declared ABI.
"""
import ctypes
import os
import re

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
HEADER = os.path.join(os.path.dirname(HERE), 'include', 'bruno_sm100.h')
LIB_PATH = os.environ.get('BRUNO_SM100_LIB') or os.path.join(HERE, 'lib', 'libbruno_sm100.so')   # override: A/B builds of the same ABI

TP, GP = 0, 1

_PROTO = re.compile(r'^\s*(const\s+char\s*\*|int|size_t|void)\s+(bruno_\w+)\s*\(([^;]*?)\)\s*;', re.M | re.S)


def _ctype(decl):
    decl = decl.strip()
    if decl in ('void', ''):
        return None
    if '*' in decl:
        return ctypes.c_void_p
    base = decl.rsplit(' ', 1)[0] if ' ' in decl else decl
    base = base.replace('const', '').strip()
    return {'int': ctypes.c_int, 'long': ctypes.c_long, 'float': ctypes.c_float, 'double': ctypes.c_double,
            'size_t': ctypes.c_size_t, 'uint64_t': ctypes.c_uint64, 'unsigned': ctypes.c_uint}[base]


def declared_prototypes(header=HEADER):
    """{name: (restype, [argtypes])} for every function the header declares."""
    src = open(header).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    out = {}
    for ret, name, args in _PROTO.findall(src):
        ret = ret.strip()
        restype = {'int': ctypes.c_int, 'size_t': ctypes.c_size_t, 'void': None}.get(ret, ctypes.c_char_p)
        argtypes = [t for t in (_ctype(a) for a in args.replace('\n', ' ').split(',')) if t is not None]
        out[name] = (restype, argtypes)
    return out


_lib = None
_protos = None


def load():
    global _lib, _protos
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError('%s is missing: build it with `python -m bruno_b200.build` '
                           '(there is no CPU fallback for the BRUNO hot path)' % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    _protos = declared_prototypes()
    for name, (restype, argtypes) in _protos.items():
        if not hasattr(lib, name) and (name.startswith('bruno_debug_') or '_set_' in name):
            continue                     # a library built with -DBRUNO_NO_DEBUG has no A/B switches
        fn = getattr(lib, name)          # AttributeError if the .so lacks a declared symbol
        fn.restype = restype
        fn.argtypes = argtypes
    _lib = lib
    return lib


def last_error():
    return load().bruno_last_error().decode()


def ptr(t, dtype=torch.float32):
    """Device pointer of a contiguous CUDA tensor (None -> NULL)."""
    if t is None:
        return None
    if not t.is_cuda:
        raise RuntimeError('bruno_b200: expected a CUDA tensor (no CPU path exists)')
    if dtype is not None and t.dtype != dtype:
        raise RuntimeError('bruno_b200: expected dtype %s, got %s' % (dtype, t.dtype))
    if not t.is_contiguous():
        raise RuntimeError('bruno_b200: expected a contiguous tensor')
    return t.data_ptr()


def stream_ptr():
    return torch.cuda.current_stream().cuda_stream


_workspaces = {}     # (device index, stream handle) -> the caller-owned GEMM workspace bound to that stream


def ensure_workspace():
    """The library never allocates: the split-K / column-reduction workspace of the dense-path GEMMs is allocated here
    (once per device and stream, zero-filled) and bound with bruno_workspace_bind.  Under stream capture nothing can be
    allocated: capture on a stream that has been warmed up (Trainer.enable_cuda_graph / BrunoModel.graphed_forward do);
    without a bound workspace the GEMMs run un-split."""
    st = torch.cuda.current_stream()
    key = (st.device.index, st.cuda_stream)
    if key in _workspaces or torch.cuda.is_current_stream_capturing():
        return
    lib = load()
    ws = torch.zeros(lib.bruno_workspace_bytes(), dtype=torch.uint8, device=st.device)
    rc = lib.bruno_workspace_bind(st.cuda_stream, ws.data_ptr(), ws.numel())
    if rc != 0:
        raise RuntimeError('bruno_workspace_bind failed (%d): %s' % (rc, last_error()))
    _workspaces[key] = ws


_ABI_DTYPES = (torch.float32, torch.uint8)      # fp32 data; uint8 = packed bf16 weights / activation slabs (opaque bytes)


def call(name, *args):
    """Invoke an int-returning entry point; tensors are passed as device pointers; appends the current stream.
    Only float32 tensors and opaque uint8 byte buffers cross the ABI: anything else (a float64 batch from a user
    iterator, an int tensor) would be reinterpreted silently, so it raises here."""
    lib = load()
    fn = getattr(lib, name)
    ensure_workspace()
    conv = []
    for a in args:
        if isinstance(a, torch.Tensor):
            if a.dtype not in _ABI_DTYPES:
                raise RuntimeError('%s: tensor of dtype %s passed to the C ABI (float32 / uint8 buffers only)' % (name, a.dtype))
            conv.append(ptr(a, None))
        else:
            conv.append(a)
    rc = fn(*conv, stream_ptr())
    if rc != 0:
        raise RuntimeError('%s failed (%d): %s' % (name, rc, last_error()))


def query(name, *args):
    """Invoke a size_t/int-returning query (no stream argument, no error code)."""
    return getattr(load(), name)(*args)
