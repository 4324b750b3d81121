"""CPU: the TensorFlow-1.x V2 checkpoint (tensor bundle) reader / writer of bruno_b200/tf_checkpoint.py -- the format of
`saver.save(sess, save_dir + '/params.ckpt')` (config_rnn/train.py:232).  TensorFlow is absent here, so the format is
pinned by its own invariants: the CRC-32C known-answer vector, LevelDB's crc mask, the table magic number, per-block and
per-tensor check-sums (a flipped byte anywhere is detected), and a write -> read round trip of a full BRUNO variable set
under the reference's variable names."""
import os
import struct

import numpy as np
import pytest

from bruno_b200 import tf_checkpoint as tfc


def test_crc32c_known_answers():
    assert tfc.crc32c(b'123456789') == 0xE3069283          # the CRC-32C (Castagnoli) check value
    assert tfc.crc32c(b'') == 0
    assert tfc.crc32c(bytes(32)) == 0x8A9136AA             # RFC 3720 B.4: 32 bytes of zeros
    assert tfc.crc32c(bytes([0xFF] * 32)) == 0x62A8AB43     # RFC 3720 B.4: 32 bytes of ones
    assert tfc.crc32c(bytes(range(32))) == 0x46DD794E       # RFC 3720 B.4: incrementing bytes
    data = os.urandom(1000)
    assert tfc.crc32c(data[300:], tfc.crc32c(data[:300])) == tfc.crc32c(data)      # streaming == one shot
    assert tfc.mask_crc(0) == 0xa282ead8


def _variables():
    rs = np.random.RandomState(0)
    v = {}
    for layer in ('Checkerboard0_1', 'Channel0_2'):
        sc = 'model/%s/function_s_t_wn/' % layer
        v[sc + 'c1/V'] = rs.randn(1, 1, 4, 64).astype(np.float32)
        v[sc + 'c1/g'] = rs.rand(64).astype(np.float32)
        for r in range(3):
            v[sc + 'c2_%d/V' % r] = rs.randn(3, 3, 64, 64).astype(np.float32)
            v[sc + 'c2_%d/b' % r] = rs.randn(64).astype(np.float32)
        v[sc + 'conv_out_scale/kernel'] = rs.randn(1, 1, 64, 4).astype(np.float32)
    v['model/even_0/function_s_t_dense_wn/d1/V'] = rs.randn(784, 512).astype(np.float32)
    v['model/prior_nu'] = rs.randn(1, 784).astype(np.float32)
    v['model/prior_nu/RMSProp'] = np.ones((1, 784), np.float32)      # optimiser slot: extra key, same bundle
    v['global_step'] = np.asarray(1234, np.int64)                   # rank-0 tensor
    return v


def test_round_trip_and_layout(tmp_path):
    prefix = os.path.join(str(tmp_path), 'params.ckpt')
    v = _variables()
    tfc.write_checkpoint(prefix, v, block_entries=4)                # several data blocks + an index block
    assert tfc.latest_checkpoint(str(tmp_path)) == prefix            # through the `checkpoint` state file
    raw = open(prefix + '.index', 'rb').read()
    assert struct.unpack('<Q', raw[-8:])[0] == tfc.MAGIC
    entries, header = tfc.list_variables(prefix)
    assert header == {'num_shards': 1, 'endianness': 0}
    assert [e[0] for e in entries] == sorted(v)                      # keys are stored sorted
    got = tfc.read_checkpoint(prefix)
    assert list(got) == sorted(v)
    for k in v:
        assert got[k].dtype == v[k].dtype and got[k].shape == v[k].shape
        assert np.array_equal(got[k], v[k]), k
    only = tfc.read_checkpoint(prefix, names={'model/prior_nu'})
    assert list(only) == ['model/prior_nu']
    # the data file is the plain concatenation of the tensors in key order
    assert os.path.getsize(prefix + '.data-00000-of-00001') == sum(a.nbytes for a in v.values())


def test_corruption_is_detected(tmp_path):
    prefix = os.path.join(str(tmp_path), 'params.ckpt')
    tfc.write_checkpoint(prefix, _variables())
    data = bytearray(open(prefix + '.data-00000-of-00001', 'rb').read())
    data[len(data) // 2] ^= 0x01
    open(prefix + '.data-00000-of-00001', 'wb').write(data)
    with pytest.raises(ValueError, match='crc32c mismatch'):
        tfc.read_checkpoint(prefix)
    assert len(tfc.read_checkpoint(prefix, verify=False)) == len(_variables())
    idx = bytearray(open(prefix + '.index', 'rb').read())
    idx[10] ^= 0x40
    open(prefix + '.index', 'wb').write(idx)
    with pytest.raises(ValueError):
        tfc.list_variables(prefix)
    bad = bytes(idx[:-8]) + struct.pack('<Q', 1)
    open(prefix + '.index', 'wb').write(bad)
    with pytest.raises(ValueError, match='magic'):
        tfc.list_variables(prefix)


def test_entry_proto_matches_hand_encoded_bytes():
    """BundleEntryProto {dtype: DT_FLOAT, shape: [3, 300], offset: 128, size: 3600, crc32c: 0x01020304} by hand:
    08 01 | 12 07 (12 02 08 03) (12 03 08 ac 02) | 20 80 01 | 28 90 1c | 35 04 03 02 01."""
    got = tfc._entry_proto(1, (3, 300), 128, 3600, 0x01020304)
    assert got == bytes.fromhex('0801' '1209' '12020803' '120308ac02' '208001' '28901c' '3504030201')
    msg = tfc._parse_proto(got)
    assert msg[1] == [1] and msg[4] == [128] and msg[5] == [3600] and msg[6] == [0x01020304]
