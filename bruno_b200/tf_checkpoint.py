"""Reader / writer of TensorFlow-1.x V2 checkpoints ("tensor bundles": `<prefix>.index` + `<prefix>.data-00000-of-00001`),
the format `tf.train.Saver().save(sess, save_dir + '/params.ckpt')` of the reference writes (config_rnn/train.py:137,232)
and `saver.restore(sess, tf.train.latest_checkpoint(save_dir))` reads (test_nll.py, test_samples.py, ...).  With it a
BRUNO checkpoint trained by the TF-1.7 reference can be scored on the B200 path, and parameters trained here can be
handed back to the reference -- no TensorFlow needed on either side of this module.

Format (restated from TensorFlow's tensor_bundle.proto, tensor_shape.proto and its LevelDB-derived table format;
TensorFlow itself is absent here, so this is validated by round trip and by the format's own check-sums, NOT against a
file written by TensorFlow):
  * `.index` is an immutable sorted string table: data blocks of prefix-compressed (key, value) entries
    [varint shared | varint non_shared | varint value_len | key suffix | value] + restart array + uint32 num_restarts;
    each block is followed by a 5-byte trailer (compression type, masked crc32c of block + type); an index block maps the
    last key of every data block to its (offset, size) handle; a 48-byte footer holds the metaindex and index handles
    and the magic number 0xdb4775248b80fb57.  TensorFlow writes bundle indices uncompressed (type 0); snappy blocks
    (type 1) are rejected with a clear error.
  * key "" -> BundleHeaderProto {1: num_shards, 2: endianness, 3: version}; every other key is a variable name ->
    BundleEntryProto {1: dtype (DT_FLOAT = 1), 2: TensorShapeProto {2: Dim {1: size}}, 3: shard_id, 4: offset, 5: size,
    6: fixed32 masked crc32c of the tensor bytes}.
  * `.data-SSSSS-of-NNNNN`: the raw little-endian tensor bytes at [offset, offset + size).
  * `checkpoint` (text): model_checkpoint_path: "<basename>" -- what tf.train.latest_checkpoint resolves.
Only float32 / float64 / int32 / int64 tensors are supported (BRUNO variables and optimiser slots are float32).
"""
import os
import struct
from collections import OrderedDict

import numpy as np

MAGIC = 0xdb4775248b80fb57
_DTYPES = {1: np.float32, 2: np.float64, 3: np.int32, 9: np.int64}
_DTYPE_IDS = {np.dtype(v): k for k, v in _DTYPES.items()}

# ---------------------------------------------------------------------------------------------------------------------
# crc32c (Castagnoli), the masked variant LevelDB / TensorFlow store
# ---------------------------------------------------------------------------------------------------------------------
_POLY = 0x82F63B78


def _make_tables():
    t0 = []
    for i in range(256):
        c = i
        for _ in range(8):
            c = (c >> 1) ^ _POLY if c & 1 else c >> 1
        t0.append(c)
    tables = [t0]
    for k in range(1, 8):
        prev = tables[-1]
        tables.append([(prev[i] >> 8) ^ t0[prev[i] & 0xFF] for i in range(256)])
    return tables


_T = _make_tables()


def crc32c(data, crc=0):
    """CRC-32C of `data` (bytes-like); slicing-by-8, ~10 MB/s in pure Python."""
    mv = memoryview(data).cast('B')
    c = crc ^ 0xFFFFFFFF
    n8 = len(mv) // 8
    if n8:
        T0, T1, T2, T3, T4, T5, T6, T7 = _T
        for lo, hi in struct.iter_unpack('<II', mv[:n8 * 8]):
            lo ^= c
            c = (T7[lo & 0xFF] ^ T6[(lo >> 8) & 0xFF] ^ T5[(lo >> 16) & 0xFF] ^ T4[lo >> 24] ^
                 T3[hi & 0xFF] ^ T2[(hi >> 8) & 0xFF] ^ T1[(hi >> 16) & 0xFF] ^ T0[hi >> 24])
    t0 = _T[0]
    for b in mv[n8 * 8:]:
        c = (c >> 8) ^ t0[(c ^ b) & 0xFF]
    return c ^ 0xFFFFFFFF


def mask_crc(crc):
    """LevelDB's crc mask: rotate right by 15 bits and add a constant (stored crcs are never crcs of crcs)."""
    return ((((crc >> 15) | (crc << 17)) & 0xFFFFFFFF) + 0xa282ead8) & 0xFFFFFFFF


# ---------------------------------------------------------------------------------------------------------------------
# varints and the few protobuf messages involved
# ---------------------------------------------------------------------------------------------------------------------
def _put_varint(out, v):
    v &= (1 << 64) - 1
    while v >= 0x80:
        out.append((v & 0x7F) | 0x80)
        v >>= 7
    out.append(v)


def _get_varint(buf, pos):
    shift = result = 0
    while True:
        b = buf[pos]
        pos += 1
        result |= (b & 0x7F) << shift
        if not b & 0x80:
            return result, pos
        shift += 7


def _parse_proto(buf):
    """{field number: [values]}; varint fields as ints, length-delimited as bytes, fixed32/64 as ints."""
    out, pos = {}, 0
    while pos < len(buf):
        key, pos = _get_varint(buf, pos)
        field, wire = key >> 3, key & 7
        if wire == 0:
            val, pos = _get_varint(buf, pos)
        elif wire == 1:
            val = struct.unpack_from('<Q', buf, pos)[0]
            pos += 8
        elif wire == 2:
            n, pos = _get_varint(buf, pos)
            val = bytes(buf[pos:pos + n])
            pos += n
        elif wire == 5:
            val = struct.unpack_from('<I', buf, pos)[0]
            pos += 4
        else:
            raise ValueError('unsupported protobuf wire type %d' % wire)
        out.setdefault(field, []).append(val)
    return out


def _entry_proto(dtype_id, shape, offset, size, crc):
    shp = bytearray()
    for d in shape:
        dim = bytearray()
        dim.append(0x08)                      # Dim.size = 1, varint
        _put_varint(dim, int(d))
        shp.append(0x12)                      # TensorShapeProto.dim = 2, length-delimited
        _put_varint(shp, len(dim))
        shp += dim
    out = bytearray()
    out.append(0x08)
    _put_varint(out, dtype_id)                # dtype = 1
    out.append(0x12)
    _put_varint(out, len(shp))                # shape = 2
    out += shp
    # shard_id = 3 is 0 (proto3 default: omitted)
    if offset:
        out.append(0x20)
        _put_varint(out, offset)              # offset = 4
    out.append(0x28)
    _put_varint(out, size)                    # size = 5
    out.append(0x35)
    out += struct.pack('<I', crc)             # crc32c = 6, fixed32
    return bytes(out)


# ---------------------------------------------------------------------------------------------------------------------
# table blocks
# ---------------------------------------------------------------------------------------------------------------------
def _read_block(f, offset, size, verify=True):
    f.seek(offset)
    raw = f.read(size + 5)
    if len(raw) != size + 5:
        raise ValueError('truncated table block at offset %d' % offset)
    block, ctype, stored = raw[:size], raw[size], struct.unpack_from('<I', raw, size + 1)[0]
    if verify and mask_crc(crc32c(raw[:size + 1])) != stored:
        raise ValueError('crc mismatch in table block at offset %d' % offset)
    if ctype == 1:
        raise ValueError('snappy-compressed index block: TensorFlow writes bundle indices uncompressed; re-save the '
                         'checkpoint (no snappy decoder in this environment)')
    if ctype != 0:
        raise ValueError('unknown block compression type %d' % ctype)
    return block


def _block_entries(block):
    num_restarts = struct.unpack_from('<I', block, len(block) - 4)[0]
    end = len(block) - 4 - 4 * num_restarts
    pos, key = 0, b''
    while pos < end:
        shared, pos = _get_varint(block, pos)
        non_shared, pos = _get_varint(block, pos)
        vlen, pos = _get_varint(block, pos)
        key = key[:shared] + bytes(block[pos:pos + non_shared])
        pos += non_shared
        yield key, bytes(block[pos:pos + vlen])
        pos += vlen


def _build_block(items, restart_interval=16):
    out, restarts, last = bytearray(), [], b''
    for i, (key, val) in enumerate(items):
        if i % restart_interval == 0:
            restarts.append(len(out))
            shared = 0
        else:
            shared = 0
            while shared < min(len(last), len(key)) and last[shared] == key[shared]:
                shared += 1
        _put_varint(out, shared)
        _put_varint(out, len(key) - shared)
        _put_varint(out, len(val))
        out += key[shared:] + val
        last = key
    if not restarts:
        restarts = [0]
    for r in restarts:
        out += struct.pack('<I', r)
    out += struct.pack('<I', len(restarts))
    return bytes(out)


def _write_block(f, block):
    off = f.tell()
    f.write(block)
    f.write(b'\x00' + struct.pack('<I', mask_crc(crc32c(block + b'\x00'))))
    return off, len(block)


def _handle(offset, size):
    out = bytearray()
    _put_varint(out, offset)
    _put_varint(out, size)
    return bytes(out)


# ---------------------------------------------------------------------------------------------------------------------
# public API
# ---------------------------------------------------------------------------------------------------------------------
def latest_checkpoint(save_dir):
    """tf.train.latest_checkpoint: the prefix named by `<save_dir>/checkpoint`, or the single `*.index` in the directory."""
    state = os.path.join(save_dir, 'checkpoint')
    if os.path.exists(state):
        for line in open(state):
            if line.startswith('model_checkpoint_path:'):
                name = line.split(':', 1)[1].strip().strip('"')
                return name if os.path.isabs(name) else os.path.join(save_dir, name)
    idx = sorted(f for f in os.listdir(save_dir) if f.endswith('.index'))
    return os.path.join(save_dir, idx[-1][:-len('.index')]) if idx else None


def list_variables(prefix, verify=True):
    """[(name, dtype, shape, shard_id, offset, size, masked_crc32c)] in key order, plus the bundle header fields."""
    with open(prefix + '.index', 'rb') as f:
        f.seek(0, 2)
        total = f.tell()
        if total < 48:
            raise ValueError('%s.index is too short to be a tensor-bundle index' % prefix)
        f.seek(total - 48)
        footer = f.read(48)
        if struct.unpack_from('<Q', footer, 40)[0] != MAGIC:
            raise ValueError('%s.index: bad magic number (not a TensorFlow V2 checkpoint index)' % prefix)
        pos = 0
        _, pos = _get_varint(footer, pos)          # metaindex handle (unused)
        _, pos = _get_varint(footer, pos)
        ioff, pos = _get_varint(footer, pos)
        isize, pos = _get_varint(footer, pos)
        entries, header = [], None
        for _, hval in _block_entries(_read_block(f, ioff, isize, verify)):
            boff, p = _get_varint(hval, 0)
            bsize, p = _get_varint(hval, p)
            for key, val in _block_entries(_read_block(f, boff, bsize, verify)):
                msg = _parse_proto(val)
                if key == b'':
                    header = {'num_shards': msg.get(1, [1])[0], 'endianness': msg.get(2, [0])[0]}
                    continue
                shape = []
                if 2 in msg:
                    for dim in _parse_proto(msg[2][0]).get(2, []):
                        shape.append(_parse_proto(dim).get(1, [0])[0])
                entries.append((key.decode(), msg.get(1, [0])[0], tuple(shape), msg.get(3, [0])[0], msg.get(4, [0])[0],
                                msg.get(5, [0])[0], msg.get(6, [None])[0]))
    if header is None:
        raise ValueError('%s.index has no bundle header entry' % prefix)
    if header['endianness'] != 0:
        raise ValueError('big-endian bundles are not supported')
    return entries, header


def read_checkpoint(prefix, names=None, verify=True):
    """{variable name: ndarray} of a TF V2 checkpoint (all variables, or those in `names`); `verify` checks every crc32c."""
    entries, header = list_variables(prefix, verify)
    out = OrderedDict()
    files = {}
    try:
        for name, dtype_id, shape, shard, offset, size, crc in entries:
            if names is not None and name not in names:
                continue
            if dtype_id not in _DTYPES:
                raise ValueError('variable %s has unsupported dtype id %d' % (name, dtype_id))
            if shard not in files:
                files[shard] = open('%s.data-%05d-of-%05d' % (prefix, shard, header['num_shards']), 'rb')
            f = files[shard]
            f.seek(offset)
            raw = f.read(size)
            if len(raw) != size:
                raise ValueError('variable %s: data file truncated' % name)
            if verify and crc is not None and mask_crc(crc32c(raw)) != crc:
                raise ValueError('variable %s: crc32c mismatch in the data file' % name)
            arr = np.frombuffer(raw, dtype=np.dtype(_DTYPES[dtype_id]).newbyteorder('<')).reshape(shape)
            out[name] = arr.astype(_DTYPES[dtype_id])
    finally:
        for f in files.values():
            f.close()
    return out


def write_checkpoint(prefix, variables, block_entries=64):
    """Writes {name: array} as a single-shard TF V2 checkpoint + the `checkpoint` state file next to it."""
    items = sorted((k.encode(), np.require(np.asarray(v), requirements='C')) for k, v in variables.items())   # keeps rank 0
    metas, offset = [], 0
    with open(prefix + '.data-00000-of-00001', 'wb') as f:
        for key, arr in items:
            if arr.dtype not in _DTYPE_IDS:
                arr = arr.astype(np.float32)
            raw = arr.astype(arr.dtype.newbyteorder('<'), copy=False).tobytes()
            f.write(raw)
            metas.append((key, _entry_proto(_DTYPE_IDS[arr.dtype], arr.shape, offset, len(raw), mask_crc(crc32c(raw)))))
            offset += len(raw)
    header = bytes([0x08, 0x01]) + bytes([0x1a, 0x02, 0x08, 0x01])      # num_shards = 1, version {producer: 1}
    entries = [(b'', header)] + metas
    with open(prefix + '.index', 'wb') as f:
        index_items = []
        for i in range(0, len(entries), block_entries):
            chunk = entries[i:i + block_entries]
            off, size = _write_block(f, _build_block(chunk))
            index_items.append((chunk[-1][0], _handle(off, size)))
        moff, msize = _write_block(f, _build_block([]))
        ioff, isize = _write_block(f, _build_block(index_items, restart_interval=1))
        footer = _handle(moff, msize) + _handle(ioff, isize)
        f.write(footer + b'\x00' * (40 - len(footer)) + struct.pack('<Q', MAGIC))
    with open(os.path.join(os.path.dirname(prefix) or '.', 'checkpoint'), 'w') as f:
        base = os.path.basename(prefix)
        f.write('model_checkpoint_path: "%s"\nall_model_checkpoint_paths: "%s"\n' % (base, base))
    return prefix
