"""CPU: TensorFlow checkpoint-V2 (tensor bundle) reader / writer.

No TensorFlow and no checkpoint file exist in the build image, so the pieces are pinned separately:
crc32c against the published check value and RFC 3720 vectors, the masking against its definition,
varint / protobuf bytes against hand-computed encodings, the LevelDB table layout (prefix compression,
restarts, footer magic) structurally; reader and writer are then tested against each other."""
import os
import struct

import numpy as np
import pytest


@pytest.fixture(scope='module')
def T(built_lib):
    from disyolo_b200 import tf_checkpoint
    return tf_checkpoint


def test_crc32c_known_answers(T):
    assert T.crc32c(b'123456789') == 0xE3069283                       # the CRC-32C check value
    assert T.crc32c(bytes(32)) == 0x8A9136AA                          # RFC 3720 B.4: 32 bytes of zeros
    assert T.crc32c(bytes([0xFF] * 32)) == 0x62A8AB43                 # RFC 3720 B.4: 32 bytes of ones
    assert T.crc32c(bytes(range(32))) == 0x46DD794E                   # RFC 3720 B.4: incrementing bytes
    big = np.random.default_rng(0).integers(0, 256, 100003, dtype=np.uint8)
    # the library's slicing-by-8 routine (>= 4 KB) agrees with the pure-Python table loop, also incrementally
    tl = T._crc_table().tolist()
    c = 0xFFFFFFFF
    for b in big.tobytes():
        c = tl[(c ^ b) & 0xFF] ^ (c >> 8)
    assert T.crc32c(big) == (~c) & 0xFFFFFFFF
    assert T.crc32c(big[5000:], T.crc32c(big[:5000])) == T.crc32c(big)


def test_crc_mask_roundtrip_and_definition(T):
    for v in (0, 1, 0xE3069283, 0xFFFFFFFF, 0x12345678):
        assert T.unmask_crc(T.mask_crc(v)) == v
    assert T.mask_crc(0) == 0xa282ead8
    assert T.mask_crc(0x00008000) == (1 + 0xa282ead8)                  # rotate right by 15


def test_varint_and_entry_bytes(T):
    assert T._put_varint(0) == b'\x00' and T._put_varint(300) == b'\xac\x02'
    assert T._get_varint(b'\xac\x02', 0) == (300, 2)
    # BundleEntryProto{dtype: DT_FLOAT, shape {dim{size:3} dim{size:300}}, offset: 16, size: 3600, crc32c: 7}
    raw = T._encode_entry(1, (3, 300), 0, 16, 3600, 7)
    assert raw == b'\x08\x01' + b'\x12\x09' + b'\x12\x02\x08\x03' + b'\x12\x03\x08\xac\x02' + b'\x20\x10' + \
        b'\x28\x90\x1c' + b'\x35' + struct.pack('<I', 7)
    e = T._parse_entry(raw)
    assert (e['dtype'], e['shape'], e['offset'], e['size'], e['crc32c']) == (1, [3, 300], 16, 3600, 7)


def test_roundtrip_network_variables(T, tmp_path):
    import disyolo_b200 as dy
    W = dy.init_weights('lively', 3)
    sub = {k: v for k, v in W.items() if any(k.startswith('yolo/convolutional%d/' % n) for n in (1, 2, 3, 59, 81, 82))}
    sub['global_step'] = np.array(500, np.int64)                       # a scalar the network does not use
    sub['yolo/convolutional81/weights/Adam'] = np.zeros((3, 3, 32, 64), np.float32)
    prefix = str(tmp_path / 'model.ckpt-500')
    T.write_checkpoint(prefix, sub, block_size=512)                    # small blocks: many data blocks + restarts
    assert os.path.exists(prefix + '.index') and os.path.exists(prefix + '.data-00000-of-00001')
    assert open(str(tmp_path / 'checkpoint')).read().startswith('model_checkpoint_path: "model.ckpt-500"')
    data = open(prefix + '.index', 'rb').read()
    assert struct.unpack('<Q', data[-8:])[0] == 0xdb4775248b80fb57     # LevelDB table magic
    header, entries = T.read_index(prefix)
    assert header['num_shards'] == 1 and sorted(entries) == sorted(sub)
    assert entries['yolo/convolutional2/weights']['shape'] == [3, 3, 32, 64]
    back = T.read_checkpoint(prefix)
    assert sorted(back) == sorted(sub)
    for k in sub:
        assert back[k].dtype == sub[k].dtype and np.array_equal(back[k], sub[k]), k
    net = T.network_variables(back)
    assert 'global_step' not in net and 'yolo/convolutional81/weights/Adam' not in net and len(net) == len(sub) - 2
    # ignore_missing_vars semantics
    some = T.read_checkpoint(prefix, names=['yolo/convolutional1/weights', 'yolo/convolutional999/weights'])
    assert list(some) == ['yolo/convolutional1/weights']


def test_corruption_is_detected(T, tmp_path):
    prefix = str(tmp_path / 'm')
    T.write_checkpoint(prefix, {'a': np.arange(100, dtype=np.float32), 'b': np.ones((4, 4), np.float32)})
    raw = bytearray(open(prefix + '.data-00000-of-00001', 'rb').read())
    raw[10] ^= 0x40
    open(prefix + '.data-00000-of-00001', 'wb').write(bytes(raw))
    with pytest.raises(ValueError, match='checksum'):
        T.read_checkpoint(prefix)
    assert np.array_equal(T.read_checkpoint(prefix, names=['b'])['b'], np.ones((4, 4), np.float32))
    idx = bytearray(open(prefix + '.index', 'rb').read())
    idx[3] ^= 0x01
    open(prefix + '.index', 'wb').write(bytes(idx))
    with pytest.raises(ValueError):
        T.read_index(prefix)
    with pytest.raises(ValueError, match='magic'):
        open(prefix + '.index', 'wb').write(b'\0' * 64)
        T.read_index(prefix)


def test_snappy_blocks_are_readable(T):
    # literal "abcd" + copy(offset 4, len 8) -> "abcdabcdabcd"
    comp = bytes([12]) + bytes([(4 - 1) << 2]) + b'abcd' + bytes([((8 - 4) << 2) | 1, 4])
    assert T._snappy_uncompress(comp) == b'abcdabcdabcd'
