"""TensorFlow Saver-V2 checkpoints ("tensor bundles") without TensorFlow.

Replaces the ``tf.train.get_checkpoint_state`` + ``Saver.restore`` step of
``luminoth/utils/predicting.py:51-63``: Luminoth's
published and user-trained checkpoints are directories holding

    checkpoint                         text: model_checkpoint_path: "model.ckpt-N"
    model.ckpt-N.index                 key -> BundleEntryProto table
    model.ckpt-N.data-00000-of-0000M   raw little-endian tensor bytes
    config.yml, classes.json           (``tools/checkpoint/__init__.py:414-526``)

Format notes (TensorFlow ``core/util/tensor_bundle`` + ``core/lib/io/table``,
which is LevelDB's table format; restated from the published format, not from
code in /root/reference -- TensorFlow is an un-vendored dependency there):

* ``.index`` is an SSTable.  Footer = last 48 bytes: metaindex BlockHandle,
  index BlockHandle (each two varint64: offset, size), zero padding to 40
  bytes, magic ``0xdb4775248b80fb57`` little-endian.
* A block is ``contents | type(1) | masked crc32c(4)``; the handle's size
  excludes the 5-byte trailer; type 0 = raw, 1 = snappy.  Contents = entries
  ``varint shared | varint non_shared | varint value_len | key_delta | value``
  followed by the restart offsets (uint32 each) and their count (uint32).
* The index block maps separator keys to data-block handles.  In the data
  blocks, key ``""`` holds ``BundleHeaderProto`` (num_shards=1, endianness=2,
  version=3) and every other key is a variable name whose value is a
  ``BundleEntryProto``: dtype=1, shape=2 (TensorShapeProto: dim=2 {size=1}),
  shard_id=3, offset=4, size=5, crc32c=6 (fixed32, masked), slices=7.
* Tensor bytes live at ``offset`` of ``<prefix>.data-%05d-of-%05d``.

PARITY UNPINNED: no TensorFlow-written checkpoint exists in /root/reference or
in this image and there is no network, so this reader is verified against the
format description through its own writer (round trips, block/CRC/snappy
cases in ``tests/test_tf_checkpoint.py``), not against a file produced by
TensorFlow.  It fails loudly (never guesses) on anything it does not
understand: partitioned variables, string tensors, big-endian bundles.
"""
import os
import re
import struct

import numpy as np

TABLE_MAGIC = 0xdb4775248b80fb57
FOOTER_LEN = 48
BLOCK_TRAILER_LEN = 5

# tensorflow/core/framework/types.proto
DTYPES = {1: np.float32, 2: np.float64, 3: np.int32, 4: np.uint8, 5: np.int16, 6: np.int8, 9: np.int64,
          10: np.bool_, 17: np.uint16, 19: np.float16, 22: np.uint32, 23: np.uint64}
DTYPE_IDS = {np.dtype(v): k for k, v in DTYPES.items()}


class CheckpointError(ValueError):
    pass


# ------------------------------------------------------------------ crc32c (Castagnoli), LevelDB masking
def _make_crc_table():
    tbl = []
    for i in range(256):
        c = i
        for _ in range(8):
            c = (c >> 1) ^ 0x82F63B78 if c & 1 else c >> 1
        tbl.append(c)
    return tbl


_CRC_TABLE = _make_crc_table()


def _crc_register(data, reg):
    tbl = _CRC_TABLE
    for b in data:
        reg = tbl[(reg ^ b) & 0xFF] ^ (reg >> 8)
    return reg


def _zero_shift_matrix(nbytes):
    """GF(2) matrix (32 column images) of "feed nbytes zero bytes" on the CRC register, by repeated squaring."""
    def apply(m, x):
        out = 0
        i = 0
        while x:
            if x & 1:
                out ^= m[i]
            x >>= 1
            i += 1
        return out
    one = [_CRC_TABLE[(1 << i) & 0xFF] ^ ((1 << i) >> 8) for i in range(32)]     # one zero byte
    result = [1 << i for i in range(32)]                                       # identity
    power = one
    while nbytes:
        if nbytes & 1:
            result = [apply(power, c) for c in result]
        power = [apply(power, c) for c in power]
        nbytes >>= 1
    return lambda x: apply(result, x)


def crc32c(data, crc=0):
    """CRC-32C.  Large buffers (weights) are cut into lanes whose registers advance together in numpy and are
    then stitched with the zero-shift operator (the register update is GF(2)-linear), ~100x the scalar loop."""
    data = bytes(data)
    lanes = 4096
    if crc != 0 or len(data) < 64 * lanes:
        return _crc_register(data, crc ^ 0xFFFFFFFF) ^ 0xFFFFFFFF
    chunk = len(data) // lanes
    body = np.frombuffer(data, np.uint8, lanes * chunk).reshape(lanes, chunk)
    tbl = np.array(_CRC_TABLE, np.uint32)
    reg = np.zeros(lanes, np.uint32)
    reg[0] = 0xFFFFFFFF
    for i in range(chunk):
        reg = tbl[(reg ^ body[:, i]) & 0xFF] ^ (reg >> np.uint32(8))
    shift = _zero_shift_matrix(chunk)
    total = int(reg[0])
    for k in range(1, lanes):
        total = shift(total) ^ int(reg[k])
    return _crc_register(data[lanes * chunk:], total) ^ 0xFFFFFFFF


def mask_crc(crc):
    return ((((crc >> 15) | (crc << 17)) & 0xFFFFFFFF) + 0xa282ead8) & 0xFFFFFFFF


def unmask_crc(masked):
    rot = (masked - 0xa282ead8) & 0xFFFFFFFF
    return ((rot >> 17) | (rot << 15)) & 0xFFFFFFFF


# ------------------------------------------------------------------ varints / tiny protobuf reader+writer
def _get_varint(buf, pos):
    result = shift = 0
    while True:
        if pos >= len(buf):
            raise CheckpointError('truncated varint')
        b = buf[pos]
        pos += 1
        result |= (b & 0x7F) << shift
        if not b & 0x80:
            return result, pos
        shift += 7
        if shift > 63:
            raise CheckpointError('varint too long')


def _put_varint(v):
    out = bytearray()
    v &= (1 << 64) - 1
    while True:
        b = v & 0x7F
        v >>= 7
        if v:
            out.append(b | 0x80)
        else:
            out.append(b)
            return bytes(out)


def _parse_proto(buf):
    """{field number: [values]} -- varint -> int, fixed32/64 -> int, length-delimited -> bytes."""
    fields = {}
    pos = 0
    while pos < len(buf):
        tag, pos = _get_varint(buf, pos)
        num, wt = tag >> 3, tag & 7
        if wt == 0:
            val, pos = _get_varint(buf, pos)
        elif wt == 1:
            val = struct.unpack_from('<Q', buf, pos)[0]; pos += 8
        elif wt == 2:
            ln, pos = _get_varint(buf, pos)
            val = bytes(buf[pos:pos + ln]); pos += ln
            if len(val) != ln:
                raise CheckpointError('truncated protobuf field')
        elif wt == 5:
            val = struct.unpack_from('<I', buf, pos)[0]; pos += 4
        else:
            raise CheckpointError('unsupported protobuf wire type %d' % wt)
        fields.setdefault(num, []).append(val)
    return fields


def _signed64(v):
    return v - (1 << 64) if v >= (1 << 63) else v


def _field(num, wt, payload):
    return _put_varint((num << 3) | wt) + payload


def _encode_shape(shape):
    out = b''
    for d in shape:
        dim = _field(1, 0, _put_varint(int(d)))
        out += _field(2, 2, _put_varint(len(dim)) + dim)
    return out


def _encode_entry(dtype_id, shape, shard, offset, size, crc):
    out = _field(1, 0, _put_varint(dtype_id))
    sh = _encode_shape(shape)
    out += _field(2, 2, _put_varint(len(sh)) + sh)
    if shard:
        out += _field(3, 0, _put_varint(shard))
    if offset:
        out += _field(4, 0, _put_varint(offset))
    out += _field(5, 0, _put_varint(size))
    out += _field(6, 5, struct.pack('<I', crc))
    return out


# ------------------------------------------------------------------ snappy (block format) decompressor
def snappy_decompress(data):
    n, pos = _get_varint(data, 0)
    out = bytearray()
    while pos < len(data):
        tag = data[pos]; pos += 1
        kind = tag & 3
        if kind == 0:                                   # literal
            ln = tag >> 2
            if ln >= 60:
                nb = ln - 59
                ln = int.from_bytes(data[pos:pos + nb], 'little'); pos += nb
            ln += 1
            out += data[pos:pos + ln]; pos += ln
        else:
            if kind == 1:
                ln = ((tag >> 2) & 7) + 4
                off = ((tag >> 5) << 8) | data[pos]; pos += 1
            elif kind == 2:
                ln = (tag >> 2) + 1
                off = data[pos] | (data[pos + 1] << 8); pos += 2
            else:
                ln = (tag >> 2) + 1
                off = int.from_bytes(data[pos:pos + 4], 'little'); pos += 4
            if off == 0 or off > len(out):
                raise CheckpointError('corrupt snappy stream (bad copy offset)')
            for _ in range(ln):                         # copies may overlap their own output
                out.append(out[-off])
    if len(out) != n:
        raise CheckpointError('corrupt snappy stream (length %d, expected %d)' % (len(out), n))
    return bytes(out)


# ------------------------------------------------------------------ SSTable reader
def _read_handle(buf, pos):
    off, pos = _get_varint(buf, pos)
    size, pos = _get_varint(buf, pos)
    return off, size, pos


def _read_block(buf, off, size, verify=True):
    end = off + size + BLOCK_TRAILER_LEN
    if end > len(buf):
        raise CheckpointError('block handle points outside the file')
    contents = bytes(buf[off:off + size])
    ctype = buf[off + size]
    if verify:
        stored = struct.unpack_from('<I', buf, off + size + 1)[0]
        actual = crc32c(bytes(buf[off:off + size + 1]))
        if unmask_crc(stored) != actual:
            raise CheckpointError('index block checksum mismatch at offset %d' % off)
    if ctype == 0:
        return contents
    if ctype == 1:
        return snappy_decompress(contents)
    raise CheckpointError('unknown block compression type %d' % ctype)


def _iter_block(block):
    if len(block) < 4:
        raise CheckpointError('block too small')
    nrestarts = struct.unpack_from('<I', block, len(block) - 4)[0]
    limit = len(block) - 4 - 4 * nrestarts
    if limit < 0:
        raise CheckpointError('corrupt block (restart array)')
    pos = 0
    key = b''
    while pos < limit:
        shared, pos = _get_varint(block, pos)
        non_shared, pos = _get_varint(block, pos)
        vlen, pos = _get_varint(block, pos)
        if shared > len(key) or pos + non_shared + vlen > limit:
            raise CheckpointError('corrupt block entry')
        key = key[:shared] + block[pos:pos + non_shared]
        pos += non_shared
        yield key, block[pos:pos + vlen]
        pos += vlen


def read_table(buf, verify=True):
    """All (key, value) pairs of an SSTable held in ``buf`` (bytes), in key order."""
    if len(buf) < FOOTER_LEN:
        raise CheckpointError('file too small to be a tensor-bundle index')
    footer = buf[len(buf) - FOOTER_LEN:]
    if struct.unpack_from('<Q', footer, 40)[0] != TABLE_MAGIC:
        raise CheckpointError('bad table magic (not a Saver-V2 .index file)')
    _, _, pos = _read_handle(footer, 0)                 # metaindex (unused)
    ioff, isize, _ = _read_handle(footer, pos)
    out = []
    for _, handle in _iter_block(_read_block(buf, ioff, isize, verify)):
        boff, bsize, _ = _read_handle(handle, 0)
        out.extend(_iter_block(_read_block(buf, boff, bsize, verify)))
    return out


# ------------------------------------------------------------------ bundle reader
class BundleReader(object):
    """``tf.train.load_checkpoint(prefix)`` for float/int tensors: names, shapes, arrays."""

    def __init__(self, prefix, verify_index=True):
        self.prefix = prefix
        index_path = prefix + '.index'
        if not os.path.exists(index_path):
            if os.path.exists(prefix):
                raise CheckpointError('%s looks like a V1 (single-file) checkpoint; only Saver-V2 bundles '
                                      '(.index + .data-*) are supported' % prefix)
            raise CheckpointError('checkpoint index %s not found' % index_path)
        with open(index_path, 'rb') as f:
            buf = f.read()
        self.entries = {}
        self.num_shards = 1
        header_seen = False
        for key, value in read_table(buf, verify_index):
            if key == b'':
                hdr = _parse_proto(value)
                self.num_shards = hdr.get(1, [1])[0]
                if hdr.get(2, [0])[0] != 0:
                    raise CheckpointError('big-endian tensor bundles are not supported')
                header_seen = True
                continue
            e = _parse_proto(value)
            if 7 in e:
                raise CheckpointError("variable '%s' is partitioned (tensor slices): not supported"
                                      % key.decode('utf-8', 'replace'))
            dtype = e.get(1, [0])[0]
            shape = []
            for sh in e.get(2, []):
                shp = _parse_proto(sh)
                if shp.get(3, [0])[0]:
                    raise CheckpointError('unknown-rank tensor in checkpoint')
                for dim in shp.get(2, []):
                    shape.append(_signed64(_parse_proto(dim).get(1, [0])[0]))
            self.entries[key.decode('utf-8')] = dict(dtype=dtype, shape=tuple(shape), shard=e.get(3, [0])[0],
                                                     offset=e.get(4, [0])[0], size=e.get(5, [0])[0],
                                                     crc=e.get(6, [0])[0])
        if not header_seen:
            raise CheckpointError('bundle header entry missing in %s' % index_path)

    def keys(self):
        return sorted(self.entries)

    def has_tensor(self, name):
        return name in self.entries

    def shape(self, name):
        return self.entries[name]['shape']

    def dtype(self, name):
        d = self.entries[name]['dtype']
        if d not in DTYPES:
            raise CheckpointError("variable '%s' has unsupported dtype enum %d" % (name, d))
        return np.dtype(DTYPES[d])

    def get_tensor(self, name, verify=False):
        if name not in self.entries:
            raise KeyError(name)
        e = self.entries[name]
        dt = self.dtype(name)
        count = int(np.prod(e['shape'], dtype=np.int64)) if e['shape'] else 1
        if count * dt.itemsize != e['size']:
            raise CheckpointError("variable '%s': %d bytes on disk, shape %s needs %d"
                                  % (name, e['size'], e['shape'], count * dt.itemsize))
        path = '%s.data-%05d-of-%05d' % (self.prefix, e['shard'], self.num_shards)
        if not os.path.exists(path):
            raise CheckpointError('checkpoint data shard %s not found' % path)
        with open(path, 'rb') as f:
            f.seek(e['offset'])
            raw = f.read(e['size'])
        if len(raw) != e['size']:
            raise CheckpointError("variable '%s': data shard truncated" % name)
        if verify and unmask_crc(e['crc']) != crc32c(raw):
            raise CheckpointError("variable '%s': tensor checksum mismatch" % name)
        return np.frombuffer(raw, dtype=dt.newbyteorder('<')).astype(dt, copy=True).reshape(e['shape'])


# ------------------------------------------------------------------ `checkpoint` state file
def get_checkpoint_state(job_dir):
    """``tf.train.get_checkpoint_state``: (latest prefix, all prefixes) or None.  Paths in the state file are
    taken relative to the directory unless absolute (the reference's ``lumi checkpoint create`` writes relative
    ones with leading indentation, ``tools/checkpoint/__init__.py:472-481``)."""
    path = os.path.join(job_dir, 'checkpoint')
    if not os.path.exists(path):
        return None
    latest, everything = None, []
    with open(path) as f:
        for line in f:
            m = re.match(r'\s*(model_checkpoint_path|all_model_checkpoint_paths)\s*:\s*"(.*)"\s*$', line)
            if not m:
                continue
            p = m.group(2)
            if not os.path.isabs(p):
                p = os.path.join(job_dir, p)
            if m.group(1) == 'model_checkpoint_path':
                latest = p
            else:
                everything.append(p)
    if latest is None and not everything:
        return None
    if not everything:
        everything = [latest]
    return (latest or everything[-1]), everything


def latest_checkpoint(job_dir):
    """The prefix ``predicting.py:54-60`` restores: the LAST of ``all_model_checkpoint_paths``."""
    state = get_checkpoint_state(job_dir)
    if not state or not state[1]:
        raise ValueError('Could not find checkpoint in {}.'.format(job_dir))
    return state[1][-1]


def load_variables(prefix, names=None, verify=False):
    """{name: ndarray} for ``names`` (default: every variable of a supported dtype)."""
    r = BundleReader(prefix)
    if names is None:
        names = [k for k in r.keys() if r.entries[k]['dtype'] in DTYPES]
    return {n: r.get_tensor(n, verify=verify) for n in names}


# ------------------------------------------------------------------ writer (single shard, uncompressed)
class _BlockBuilder(object):
    def __init__(self, restart_interval=16):
        self.buf = bytearray()
        self.restarts = [0]
        self.counter = 0
        self.last_key = b''
        self.interval = restart_interval

    def add(self, key, value):
        shared = 0
        if self.counter < self.interval:
            m = min(len(key), len(self.last_key))
            while shared < m and key[shared] == self.last_key[shared]:
                shared += 1
        else:
            self.restarts.append(len(self.buf))
            self.counter = 0
        self.buf += _put_varint(shared) + _put_varint(len(key) - shared) + _put_varint(len(value))
        self.buf += key[shared:] + value
        self.last_key = key
        self.counter += 1

    def finish(self):
        out = bytes(self.buf)
        for r in self.restarts:
            out += struct.pack('<I', r)
        return out + struct.pack('<I', len(self.restarts))

    def size(self):
        return len(self.buf) + 4 * len(self.restarts) + 4


def _emit_block(out, contents):
    off = len(out)
    out += contents + b'\x00'
    out += struct.pack('<I', mask_crc(crc32c(contents + b'\x00')))
    return off, len(contents)


def write_bundle(prefix, tensors, block_size=4096):
    """Write {name: ndarray} as a one-shard Saver-V2 bundle (used by the tests and to export synthetic or
    converted weights in the format the reference's tooling expects), plus nothing else: the caller writes the
    ``checkpoint`` state file with :func:`write_checkpoint_state`."""
    data = bytearray()
    entries = []
    for name in sorted(tensors, key=lambda s: s.encode('utf-8')):
        arr = np.asarray(tensors[name], order='C')
        if arr.dtype not in DTYPE_IDS:
            raise CheckpointError("cannot store dtype %s ('%s')" % (arr.dtype, name))
        raw = arr.astype(arr.dtype.newbyteorder('<'), copy=False).tobytes()
        entries.append((name.encode('utf-8'),
                        _encode_entry(DTYPE_IDS[arr.dtype], arr.shape, 0, len(data), len(raw), mask_crc(crc32c(raw)))))
        data += raw
    version = _field(1, 0, _put_varint(1))                                   # VersionDef.producer = 1
    header = _field(1, 0, _put_varint(1)) + _field(3, 2, _put_varint(len(version)) + version)
    items = [(b'', header)] + entries

    out = bytearray()
    index = _BlockBuilder(restart_interval=1)
    blk = _BlockBuilder()
    pending = None
    for key, value in items:
        if pending is not None:                       # separator for the finished block: its last key
            index.add(pending[0], _put_varint(pending[1]) + _put_varint(pending[2]))
            pending = None
        blk.add(key, value)
        if blk.size() >= block_size:
            off, size = _emit_block(out, blk.finish())
            pending = (blk.last_key, off, size)
            blk = _BlockBuilder()
    if blk.counter or not out:
        off, size = _emit_block(out, blk.finish())
        pending = (blk.last_key, off, size)
    if pending is not None:
        index.add(pending[0], _put_varint(pending[1]) + _put_varint(pending[2]))
    moff, msize = _emit_block(out, _BlockBuilder().finish())          # empty metaindex
    ioff, isize = _emit_block(out, index.finish())
    footer = _put_varint(moff) + _put_varint(msize) + _put_varint(ioff) + _put_varint(isize)
    footer += b'\x00' * (40 - len(footer)) + struct.pack('<Q', TABLE_MAGIC)
    out += footer
    os.makedirs(os.path.dirname(os.path.abspath(prefix)), exist_ok=True)
    with open(prefix + '.index', 'wb') as f:
        f.write(bytes(out))
    with open(prefix + '.data-00000-of-00001', 'wb') as f:
        f.write(bytes(data))


def write_checkpoint_state(job_dir, prefix_basename):
    with open(os.path.join(job_dir, 'checkpoint'), 'w') as f:
        f.write('model_checkpoint_path: "{0}"\nall_model_checkpoint_paths: "{0}"\n'.format(prefix_basename))
