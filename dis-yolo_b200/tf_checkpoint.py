"""TensorFlow checkpoint V2 ("tensor bundle") reader / writer without TensorFlow  (SURVEY section 8 row f-1).

Replaces, for the variables the hot path needs, `tf.train.Saver.restore` / `slim.assign_from_checkpoint_fn`
(train_yolo3_mask.py:85-111, calculate_test_map.py:184-185) and `Saver.save` (train_yolo3_mask.py:221-226):
the reference's `yolov3_3class_coco.ckpt` and `model.ckpt-<step>` are bundles

    <prefix>.index                 an SSTable: key = variable name, value = BundleEntryProto
                                   (key "" = BundleHeaderProto)
    <prefix>.data-0000N-of-0000M   raw little-endian tensor bytes at (shard_id, offset, size)

Format restated from TensorFlow's tensor_bundle / lib/io/table sources (LevelDB table format):
  data / index blocks   entries [varint32 shared][varint32 non_shared][varint32 value_len][key suffix][value],
                        then uint32 restart offsets, uint32 num_restarts; block trailer = 1 type byte
                        (0 raw, 1 snappy) + masked crc32c of (block + type)
  footer (48 bytes)     metaindex BlockHandle, index BlockHandle (varint64 offset, size), zero padding,
                        magic 0xdb4775248b80fb57
  BundleEntryProto      1 dtype, 2 shape {2 dim {1 size}}, 3 shard_id, 4 offset, 5 size, 6 crc32c (fixed32,
                        masked crc32c of the tensor bytes), 7 slices (partitioned variables: not supported)

PARITY STATUS: unpinned -- no TensorFlow and no checkpoint file exist in the build image (the reference
ships none: its weights are Google-Drive links).  Pinned pieces: crc32c and its masking against the
published known answers, varint / protobuf wire encoding against hand-computed bytes; the reader and
the writer are tested against each other (tests/test_tf_checkpoint_cpu.py).
"""
import os
import struct

import numpy as np

MAGIC = 0xdb4775248b80fb57
_DTYPES = {1: np.float32, 2: np.float64, 3: np.int32, 9: np.int64, 4: np.uint8, 6: np.int8, 10: np.bool_,
           19: np.float16}
_DTYPE_CODE = {np.dtype(v): k for k, v in _DTYPES.items()}

# ---- crc32c (Castagnoli), table driven ---------------------------------------------------------
_CRC_TABLE = None


def _crc_table():
    global _CRC_TABLE
    if _CRC_TABLE is None:
        t = np.zeros(256, np.uint32)
        for i in range(256):
            c = i
            for _ in range(8):
                c = (c >> 1) ^ 0x82F63B78 if c & 1 else c >> 1
            t[i] = c
        _CRC_TABLE = t
    return _CRC_TABLE


def crc32c(data, crc=0):
    """CRC-32C of bytes-like `data`: the library's slicing-by-8 routine (dy_crc32c) for anything large,
    a table-driven pure-Python loop otherwise (and when the library has not been built)."""
    mv0 = memoryview(data).cast('B')
    if len(mv0) >= 4096:
        try:
            import ctypes
            from . import _lib
            buf = (ctypes.c_char * len(mv0)).from_buffer_copy(mv0) if mv0.readonly else \
                (ctypes.c_char * len(mv0)).from_buffer(mv0)
            return int(_lib.lib().dy_crc32c(ctypes.cast(buf, ctypes.c_void_p), len(mv0), crc))
        except (RuntimeError, OSError):
            pass
    t = _crc_table()
    c = (~crc) & 0xFFFFFFFF
    mv = memoryview(data).cast('B')
    tl = t.tolist()
    for b in mv:
        c = tl[(c ^ b) & 0xFF] ^ (c >> 8)
    return (~c) & 0xFFFFFFFF


def mask_crc(crc):
    return ((((crc >> 15) | (crc << 17)) & 0xFFFFFFFF) + 0xa282ead8) & 0xFFFFFFFF


def unmask_crc(masked):
    rot = (masked - 0xa282ead8) & 0xFFFFFFFF
    return ((rot >> 17) | (rot << 15)) & 0xFFFFFFFF


# ---- varints / protobuf wire format ---------------------------------------------------------------
def _put_varint(v):
    out = bytearray()
    v &= (1 << 64) - 1
    while v >= 0x80:
        out.append((v & 0x7F) | 0x80)
        v >>= 7
    out.append(v)
    return bytes(out)


def _get_varint(buf, pos):
    shift, v = 0, 0
    while True:
        b = buf[pos]
        pos += 1
        v |= (b & 0x7F) << shift
        if not b & 0x80:
            return v, pos
        shift += 7


def _pb_fields(buf):
    """Yield (field_number, wire_type, value) of one protobuf message."""
    pos, n = 0, len(buf)
    while pos < n:
        key, pos = _get_varint(buf, pos)
        f, wt = key >> 3, key & 7
        if wt == 0:
            v, pos = _get_varint(buf, pos)
        elif wt == 1:
            v = struct.unpack_from('<Q', buf, pos)[0]
            pos += 8
        elif wt == 2:
            ln, pos = _get_varint(buf, pos)
            v = bytes(buf[pos:pos + ln])
            pos += ln
        elif wt == 5:
            v = struct.unpack_from('<I', buf, pos)[0]
            pos += 4
        else:
            raise ValueError('unsupported protobuf wire type %d' % wt)
        yield f, wt, v


def _parse_entry(buf):
    e = dict(dtype=0, shape=[], shard_id=0, offset=0, size=0, crc32c=0, slices=False)
    for f, wt, v in _pb_fields(buf):
        if f == 1:
            e['dtype'] = v
        elif f == 2:
            for f2, _, v2 in _pb_fields(v):
                if f2 == 2:                       # TensorShapeProto.dim
                    size = 0
                    for f3, _, v3 in _pb_fields(v2):
                        if f3 == 1:
                            size = v3 if v3 < (1 << 63) else v3 - (1 << 64)
                    e['shape'].append(size)
        elif f == 3:
            e['shard_id'] = v
        elif f == 4:
            e['offset'] = v
        elif f == 5:
            e['size'] = v
        elif f == 6:
            e['crc32c'] = v
        elif f == 7:
            e['slices'] = True
    return e


def _encode_entry(dtype_code, shape, shard_id, offset, size, crc):
    dims = b''.join(b'\x12' + _put_varint(len(d)) + d for d in (b'\x08' + _put_varint(int(s)) for s in shape))
    out = b'\x08' + _put_varint(dtype_code) + b'\x12' + _put_varint(len(dims)) + dims
    if shard_id:
        out += b'\x18' + _put_varint(shard_id)
    if offset:
        out += b'\x20' + _put_varint(offset)
    out += b'\x28' + _put_varint(size) + b'\x35' + struct.pack('<I', crc)
    return out


# ---- snappy (raw format) decompression, for index files written with compressed blocks ---------------
def _snappy_uncompress(buf):
    n, pos = _get_varint(buf, 0)
    out = bytearray()
    while pos < len(buf):
        tag = buf[pos]
        pos += 1
        kind = tag & 3
        if kind == 0:
            ln = tag >> 2
            if ln >= 60:
                nb = ln - 59
                ln = int.from_bytes(buf[pos:pos + nb], 'little')
                pos += nb
            ln += 1
            out += buf[pos:pos + ln]
            pos += ln
            continue
        if kind == 1:
            ln = ((tag >> 2) & 7) + 4
            off = ((tag >> 5) << 8) | buf[pos]
            pos += 1
        elif kind == 2:
            ln = (tag >> 2) + 1
            off = buf[pos] | (buf[pos + 1] << 8)
            pos += 2
        else:
            ln = (tag >> 2) + 1
            off = int.from_bytes(buf[pos:pos + 4], 'little')
            pos += 4
        for _ in range(ln):                      # copies may overlap their own output
            out.append(out[-off])
    if len(out) != n:
        raise ValueError('corrupt snappy block')
    return bytes(out)


# ---- table reading ------------------------------------------------------------------------------------
def _read_block(data, offset, size, verify=True):
    block = data[offset:offset + size]
    ctype = data[offset + size]
    if verify:
        stored = struct.unpack_from('<I', data, offset + size + 1)[0]
        if unmask_crc(stored) != crc32c(data[offset:offset + size + 1]):
            raise ValueError('checkpoint index: block checksum mismatch at offset %d' % offset)
    if ctype == 1:
        block = _snappy_uncompress(block)
    elif ctype != 0:
        raise ValueError('checkpoint index: unknown block compression %d' % ctype)
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


def read_index(prefix, verify=True):
    """-> (header dict, {name: entry dict}) of <prefix>.index."""
    with open(prefix + '.index', 'rb') as f:
        data = f.read()
    if len(data) < 48 or struct.unpack_from('<Q', data, len(data) - 8)[0] != MAGIC:
        raise ValueError('%s.index is not a TensorFlow checkpoint-V2 index (bad magic)' % prefix)
    foot = data[-48:]
    _, p = _get_varint(foot, 0)
    _, p = _get_varint(foot, p)                  # metaindex handle (unused)
    ioff, p = _get_varint(foot, p)
    isize, p = _get_varint(foot, p)
    entries, header = {}, dict(num_shards=1, endianness=0)
    for _, handle in _block_entries(_read_block(data, ioff, isize, verify)):
        boff, q = _get_varint(handle, 0)
        bsize, q = _get_varint(handle, q)
        for key, value in _block_entries(_read_block(data, boff, bsize, verify)):
            if key == b'':
                for f_, _, v in _pb_fields(value):
                    if f_ == 1:
                        header['num_shards'] = v
                    elif f_ == 2:
                        header['endianness'] = v
            else:
                entries[key.decode('utf-8')] = _parse_entry(value)
    if header['endianness'] != 0:
        raise ValueError('big-endian checkpoints are not supported')
    return header, entries


def read_checkpoint(prefix, names=None, verify=True):
    """{variable name: ndarray} of a checkpoint-V2 bundle.  `names`: optional iterable restricting what is
    read (missing names are skipped, like slim's ignore_missing_vars=True, train_yolo3_mask.py:106)."""
    header, entries = read_index(prefix, verify)
    want = set(entries) if names is None else (set(names) & set(entries))
    out, files = {}, {}
    try:
        for name in sorted(want):
            e = entries[name]
            if e['slices']:
                raise ValueError('partitioned variable %s is not supported' % name)
            if e['dtype'] not in _DTYPES:
                continue                          # strings / resources: not tensors the network uses
            path = '%s.data-%05d-of-%05d' % (prefix, e['shard_id'], header['num_shards'])
            f = files.get(path)
            if f is None:
                f = files[path] = open(path, 'rb')
            f.seek(e['offset'])
            raw = f.read(e['size'])
            if len(raw) != e['size']:
                raise ValueError('%s: truncated data file' % name)
            if verify and e['crc32c'] and unmask_crc(e['crc32c']) != crc32c(raw):
                raise ValueError('%s: tensor checksum mismatch' % name)
            out[name] = np.frombuffer(raw, dtype=_DTYPES[e['dtype']]).reshape(e['shape']).copy()
    finally:
        for f in files.values():
            f.close()
    return out


# ---- writing -------------------------------------------------------------------------------------------
class _BlockBuilder(object):
    def __init__(self, restart_interval=16):
        self.buf, self.restarts, self.count, self.last, self.ri = bytearray(), [0], 0, b'', restart_interval

    def add(self, key, value):
        shared = 0
        if self.count % self.ri == 0 and self.count:
            self.restarts.append(len(self.buf))
        elif self.count:
            m = min(len(key), len(self.last))
            while shared < m and key[shared] == self.last[shared]:
                shared += 1
        self.buf += _put_varint(shared) + _put_varint(len(key) - shared) + _put_varint(len(value))
        self.buf += key[shared:] + value
        self.last, self.count = key, self.count + 1

    def finish(self):
        return bytes(self.buf) + b''.join(struct.pack('<I', r) for r in self.restarts) + \
            struct.pack('<I', len(self.restarts))


def _emit_block(out, block):
    off = len(out)
    out += block + b'\x00'
    out += struct.pack('<I', mask_crc(crc32c(block + b'\x00')))
    return _put_varint(off) + _put_varint(len(block))


def write_checkpoint(prefix, tensors, block_size=4096, with_crc=True):
    """Write {name: ndarray} as a single-shard checkpoint-V2 bundle (+ the `checkpoint` state file that
    tf.train.latest_checkpoint reads) -- the counterpart of Saver.save, train_yolo3_mask.py:221-226."""
    d = os.path.dirname(os.path.abspath(prefix))
    os.makedirs(d, exist_ok=True)
    items = sorted((k.encode('utf-8'), np.asarray(v).copy(order='C')) for k, v in tensors.items())   # keeps 0-d
    records, offset = [], 0
    with open(prefix + '.data-00000-of-00001', 'wb') as f:
        for key, a in items:
            if a.dtype not in _DTYPE_CODE:
                raise ValueError('unsupported dtype %s' % a.dtype)
            raw = a.tobytes()
            f.write(raw)
            crc = mask_crc(crc32c(raw)) if with_crc else 0
            records.append((key, _encode_entry(_DTYPE_CODE[a.dtype], a.shape, 0, offset, len(raw), crc)))
            offset += len(raw)
    header = b'\x08\x01' + b'\x1a\x02\x08\x01'             # num_shards=1, (endianness LITTLE=default), version{producer=1}
    out = bytearray()
    index = _BlockBuilder(restart_interval=1)
    blk, last_key = _BlockBuilder(), b''
    for key, value in [(b'', header)] + records:
        blk.add(key, value)
        last_key = key
        if len(blk.buf) >= block_size:
            index.add(last_key, _emit_block(out, blk.finish()))
            blk = _BlockBuilder()
    if blk.count:
        index.add(last_key, _emit_block(out, blk.finish()))
    meta = _emit_block(out, _BlockBuilder().finish())
    idx = _emit_block(out, index.finish())
    foot = meta + idx
    out += foot + b'\x00' * (40 - len(foot)) + struct.pack('<Q', MAGIC)
    with open(prefix + '.index', 'wb') as f:
        f.write(bytes(out))
    with open(os.path.join(d, 'checkpoint'), 'w') as f:
        base = os.path.basename(prefix)
        f.write('model_checkpoint_path: "%s"\nall_model_checkpoint_paths: "%s"\n' % (base, base))


def is_checkpoint_prefix(path):
    return isinstance(path, str) and os.path.exists(path + '.index')


def network_variables(tensors):
    """Keep what the network uses: yolo/convolutional{N}/... (drops Adam slots, global_step, ...)."""
    keep = ('/weights', '/biases', '/BatchNorm/gamma', '/BatchNorm/beta', '/BatchNorm/moving_mean',
            '/BatchNorm/moving_variance')
    return {k: v for k, v in tensors.items() if k.startswith('yolo/convolutional') and k.endswith(keep)}
