"""Saver-V2 tensor-bundle reader/writer (replaces predicting.py:51-63's Saver.restore).

No TensorFlow-written checkpoint is available offline, so these tests pin the implementation on the published
format through independent constructions: a hand-assembled table, known CRC-32C / snappy vectors, and round
trips through the writer."""
import os
import struct

import numpy as np
import pytest

from luminoth_b200 import tf_checkpoint as tfc


def test_crc32c_known_vectors_and_mask_roundtrip():
    # RFC 3720 B.4 test vectors for CRC-32C (Castagnoli)
    assert tfc.crc32c(b'') == 0
    assert tfc.crc32c(b'123456789') == 0xE3069283
    assert tfc.crc32c(bytes(32)) == 0x8A9136AA
    assert tfc.crc32c(bytes([0xFF] * 32)) == 0x62A8AB43
    assert tfc.crc32c(bytes(range(32))) == 0x46DD794E
    for v in (0, 1, 0xE3069283, 0xFFFFFFFF):
        assert tfc.unmask_crc(tfc.mask_crc(v)) == v
    assert tfc.mask_crc(0) == 0xa282ead8                      # LevelDB: rotate right 15, add kMaskDelta


def test_snappy_literals_and_overlapping_copies():
    # preamble varint(len) | literal "abcd" | copy(offset 4, len 8, 1-byte-offset form) -> "abcdabcdabcd"
    stream = bytes([12, (4 - 1) << 2]) + b'abcd' + bytes([((8 - 4) << 2) | 1, 4])
    assert tfc.snappy_decompress(stream) == b'abcdabcdabcd'
    # long literal (length byte form 60) and a 2-byte-offset copy
    lit = bytes(range(100))
    stream = bytes([110, 60 << 2, 99]) + lit + bytes([((10 - 1) << 2) | 2, 100, 0])
    assert tfc.snappy_decompress(stream) == lit + lit[:10]
    with pytest.raises(tfc.CheckpointError):
        tfc.snappy_decompress(bytes([5, (4 - 1) << 2]) + b'abcd')         # declared length mismatch


def _hand_table(entries, compress=False):
    """An SSTable assembled here from the format description, independently of tfc's writer: one data block,
    restart at every entry (no prefix sharing), optional snappy (literal-only) block."""
    def varint(v):
        out = b''
        while v >= 0x80:
            out += bytes([v & 0x7F | 0x80]); v >>= 7
        return out + bytes([v])

    def block(kvs):
        body, restarts = b'', []
        for k, v in kvs:
            restarts.append(len(body))
            body += varint(0) + varint(len(k)) + varint(len(v)) + k + v
        for r in restarts or [0]:
            body += struct.pack('<I', r)
        return body + struct.pack('<I', len(restarts) or 1)

    def emit(buf, contents, ctype=0):
        if ctype == 1:
            n = len(contents)
            assert 60 < n <= 256                      # literal with a one-byte length (tag 60)
            contents = varint(n) + bytes([60 << 2, n - 1]) + contents
        off = len(buf)
        buf += contents + bytes([ctype])
        buf += struct.pack('<I', tfc.mask_crc(tfc.crc32c(contents + bytes([ctype]))))
        return off, len(contents)

    buf = bytearray()
    doff, dsize = emit(buf, block(entries), 1 if compress else 0)
    moff, msize = emit(buf, block([]))
    ioff, isize = emit(buf, block([(entries[-1][0], varint(doff) + varint(dsize))]))
    footer = varint(moff) + varint(msize) + varint(ioff) + varint(isize)
    buf += footer + bytes(40 - len(footer)) + struct.pack('<Q', 0xdb4775248b80fb57)
    return bytes(buf)


def test_reader_on_hand_assembled_bundle(tmp_path):
    """BundleEntryProto bytes written out by hand: dtype=DT_FLOAT(1), shape {dim{size:2} dim{size:3}}, offset 8,
    size 24, plus an int64 scalar (global_step) at offset 0."""
    w = np.arange(6, dtype=np.float32).reshape(2, 3) - 2.5
    step = np.array(1234567, np.int64)
    data = step.tobytes() + w.tobytes()
    crc_w = tfc.mask_crc(tfc.crc32c(w.tobytes())); crc_s = tfc.mask_crc(tfc.crc32c(step.tobytes()))
    entry_w = (b'\x08\x01' + b'\x12\x08' + b'\x12\x02\x08\x02' + b'\x12\x02\x08\x03' + b'\x20\x08' + b'\x28\x18'
               + b'\x35' + struct.pack('<I', crc_w))
    entry_s = b'\x08\x09' + b'\x12\x00' + b'\x28\x08' + b'\x35' + struct.pack('<I', crc_s)
    header = b'\x08\x01' + b'\x1a\x02\x08\x01'               # num_shards=1, version{producer=1}
    for compress in (False, True):
        prefix = str(tmp_path / ('model.ckpt-%d' % compress))
        table = _hand_table([(b'', header), (b'global_step', entry_s), (b'rpn/conv/w', entry_w)], compress)
        with open(prefix + '.index', 'wb') as f:
            f.write(table)
        with open(prefix + '.data-00000-of-00001', 'wb') as f:
            f.write(data)
        r = tfc.BundleReader(prefix)
        assert r.keys() == ['global_step', 'rpn/conv/w']
        assert r.shape('rpn/conv/w') == (2, 3) and r.shape('global_step') == ()
        np.testing.assert_array_equal(r.get_tensor('rpn/conv/w', verify=True), w)
        assert r.get_tensor('global_step', verify=True) == 1234567
    # a flipped byte in the index is caught by the block checksum
    bad = bytearray(table); bad[3] ^= 0x40
    with open(prefix + '.index', 'wb') as f:
        f.write(bytes(bad))
    with pytest.raises(tfc.CheckpointError):
        tfc.BundleReader(prefix)


def test_writer_reader_roundtrip_many_blocks(tmp_path):
    """Enough variables for several 4 KB data blocks with prefix-compressed keys (ResNet-style scopes)."""
    rng = np.random.default_rng(0)
    tensors = {'global_step': np.array(7, np.int64)}
    for b in range(1, 4):
        for u in range(1, 24):
            for c in ('conv1', 'conv2', 'conv3'):
                base = 'truncated_base_network/resnet_v1_101/block%d/unit_%d/bottleneck_v1/%s' % (b, u, c)
                tensors[base + '/weights'] = rng.standard_normal((1, 1, 4, 3)).astype(np.float32)
                for v in ('gamma', 'beta', 'moving_mean', 'moving_variance'):
                    tensors[base + '/BatchNorm/' + v] = rng.standard_normal((3,)).astype(np.float32)
    tensors['half'] = rng.standard_normal((5,)).astype(np.float16)
    tensors['flags'] = np.array([True, False, True])
    prefix = str(tmp_path / 'run' / 'model.ckpt-7')
    tfc.write_bundle(prefix, tensors)
    assert os.path.getsize(prefix + '.index') > 3 * 4096
    r = tfc.BundleReader(prefix)
    assert r.keys() == sorted(tensors, key=lambda s: s.encode())
    for k, v in tensors.items():
        got = r.get_tensor(k, verify=True)
        assert got.dtype == v.dtype and got.shape == v.shape
        np.testing.assert_array_equal(got, v)
    some = tfc.load_variables(prefix, ['half', 'flags'])
    assert set(some) == {'half', 'flags'}


def test_checkpoint_state_file(tmp_path):
    d = str(tmp_path)
    assert tfc.get_checkpoint_state(d) is None
    with pytest.raises(ValueError, match='Could not find checkpoint'):
        tfc.latest_checkpoint(d)
    # the indented form written by `lumi checkpoint create` (tools/checkpoint/__init__.py:474-481)
    with open(os.path.join(d, 'checkpoint'), 'w') as f:
        f.write('\n            model_checkpoint_path: "model.ckpt-90000"\n'
                '            all_model_checkpoint_paths: "model.ckpt-80000"\n'
                '            all_model_checkpoint_paths: "model.ckpt-90000"\n            ')
    latest, everything = tfc.get_checkpoint_state(d)
    assert latest == os.path.join(d, 'model.ckpt-90000')
    assert everything == [os.path.join(d, 'model.ckpt-80000'), os.path.join(d, 'model.ckpt-90000')]
    assert tfc.latest_checkpoint(d) == os.path.join(d, 'model.ckpt-90000')      # predicting.py:60 takes [-1]
    tfc.write_checkpoint_state(d, '/abs/elsewhere/model.ckpt-1')
    assert tfc.latest_checkpoint(d) == '/abs/elsewhere/model.ckpt-1'


def test_reader_refuses_what_it_does_not_understand(tmp_path):
    prefix = str(tmp_path / 'model.ckpt-1')
    with pytest.raises(tfc.CheckpointError, match='not found'):
        tfc.BundleReader(prefix)
    open(prefix, 'wb').write(b'v1 checkpoint bytes')
    with pytest.raises(tfc.CheckpointError, match='V1'):
        tfc.BundleReader(prefix)
    os.remove(prefix)
    open(prefix + '.index', 'wb').write(b'\x00' * 64)
    with pytest.raises(tfc.CheckpointError, match='magic'):
        tfc.BundleReader(prefix)
    # a partitioned variable (slices field 7 present)
    header = b'\x08\x01'
    entry = b'\x08\x01' + b'\x12\x04\x12\x02\x08\x02' + b'\x28\x08' + b'\x3a\x00'
    open(prefix + '.index', 'wb').write(_hand_table([(b'', header), (b'part', entry)]))
    with pytest.raises(tfc.CheckpointError, match='partitioned'):
        tfc.BundleReader(prefix)
    # size / shape disagreement
    entry = b'\x08\x01' + b'\x12\x04\x12\x02\x08\x02' + b'\x28\x04'
    open(prefix + '.index', 'wb').write(_hand_table([(b'', header), (b'w', entry)]))
    open(prefix + '.data-00000-of-00001', 'wb').write(b'\x00' * 8)
    with pytest.raises(tfc.CheckpointError, match='bytes on disk'):
        tfc.BundleReader(prefix).get_tensor('w')


def test_crc32c_lane_parallel_path_equals_scalar():
    data = np.random.default_rng(3).integers(0, 256, 300_007, dtype=np.uint8).tobytes()
    assert tfc.crc32c(data) == (tfc._crc_register(data, 0xFFFFFFFF) ^ 0xFFFFFFFF)


class _FakeEngine(object):
    def __init__(self, specs):
        self._specs = specs

    def weight_specs(self):
        return self._specs


def test_load_checkpoint_weights_selects_model_variables(tmp_path):
    """predicting.py:51-63: variables are restored by name; slots the inference graph does not own are ignored,
    a missing or mis-shaped model variable is a ValueError."""
    from luminoth_b200.predicting import load_checkpoint_weights
    d = str(tmp_path)
    rng = np.random.default_rng(1)
    tensors = {'fasterrcnn/rpn/conv/w': rng.standard_normal((3, 3, 8, 4)).astype(np.float32),
               'fasterrcnn/rpn/conv/b': rng.standard_normal((4,)).astype(np.float32),
               'fasterrcnn/rpn/conv/w/Momentum': np.zeros((3, 3, 8, 4), np.float32),
               'global_step': np.array(10, np.int64)}
    tfc.write_bundle(os.path.join(d, 'model.ckpt-10'), tensors)
    tfc.write_checkpoint_state(d, 'model.ckpt-10')
    eng = _FakeEngine([('fasterrcnn/rpn/conv/w', (3, 3, 8, 4)), ('fasterrcnn/rpn/conv/b', (4,))])
    got = load_checkpoint_weights(eng, d)
    assert set(got) == {'fasterrcnn/rpn/conv/w', 'fasterrcnn/rpn/conv/b'}
    np.testing.assert_array_equal(got['fasterrcnn/rpn/conv/w'], tensors['fasterrcnn/rpn/conv/w'])
    with pytest.raises(ValueError, match='lacks 1 model variables'):
        load_checkpoint_weights(_FakeEngine([('fasterrcnn/rpn/cls_conv/w', (1, 1, 4, 2))]), d)
    with pytest.raises(ValueError, match='has shape'):
        load_checkpoint_weights(_FakeEngine([('fasterrcnn/rpn/conv/b', (8,))]), d)
    with pytest.raises(ValueError, match='Could not find checkpoint'):
        load_checkpoint_weights(eng, os.path.join(d, 'nothing_here'))
