"""TF V2 checkpoint bundle reader/writer (reference: tf.train.Saver at main.py:186-191, 211, 288).

TensorFlow is not installable here and the reference ships no checkpoint, so the format is pinned by the published
constants and known answers of its building blocks: the CRC-32C vectors of RFC 3720 B.4 (the ones leveldb's and TF's
crc32c_test use), the crc mask of leveldb/TF, the table magic number, hand-assembled protobuf / block bytes, and a
real snappy compressor (pyarrow's) for the read-side decompressor."""
import os
import struct

import numpy as np
import pytest

from vae_captioning_b200 import tf_bundle as tb


def test_crc32c_known_answers():
    assert tb._crc32c_py(b"123456789") == 0xE3069283
    assert tb._crc32c_py(bytes(32)) == 0x8A9136AA
    assert tb._crc32c_py(b"\xff" * 32) == 0x62A8AB43
    assert tb._crc32c_py(bytes(range(32))) == 0x46DD794E
    assert tb._crc32c_py(bytes(range(31, -1, -1))) == 0x113FDB5C
    # extend: crc(a + b) == crc(b, seed=crc(a))
    assert tb._crc32c_py(b"world", tb._crc32c_py(b"hello ")) == tb._crc32c_py(b"hello world")
    # mask is a rotate + constant and must invert (leveldb crc32c.h)
    c = tb._crc32c_py(b"foo")
    assert tb.mask_crc(c) != c and tb.unmask_crc(tb.mask_crc(c)) == c
    assert tb.mask_crc(0) == 0xA282EAD8


def test_crc32c_native_helper_matches_table_version():
    """vc_crc32c (slicing-by-8 in libvaecap.so, host only) against the byte-at-a-time table on odd sizes / alignments."""
    from vae_captioning_b200 import build
    build.build()
    rng = np.random.Generator(np.random.PCG64(3))
    raw = rng.integers(0, 256, size=70001, dtype=np.uint8)
    for off, n in ((0, 70001), (1, 65536), (3, 4097), (7, 8192 + 5)):
        a = raw[off:off + n]
        assert tb.crc32c(a) == tb._crc32c_py(a.tobytes())
    assert tb.crc32c(raw[:5000], 0x1234) == tb._crc32c_py(raw[:5000].tobytes(), 0x1234)
    big = np.zeros(32, np.uint8)
    assert tb.crc32c(big) == 0x8A9136AA


def test_snappy_reader_against_a_real_compressor():
    pa = pytest.importorskip("pyarrow")
    rng = np.random.Generator(np.random.PCG64(4))
    cases = [b"", b"a", b"abc" * 1000, bytes(rng.integers(0, 4, size=20000, dtype=np.uint8)),
             bytes(rng.integers(0, 256, size=70000, dtype=np.uint8)),
             b"encoder/gmm_ll_12/dense_1/kernel" * 300 + bytes(range(256)) * 40]
    for raw in cases:
        comp = pa.compress(raw, codec="snappy", asbytes=True)
        assert tb.snappy_uncompress(comp) == raw
    with pytest.raises(tb.BundleError):
        tb.snappy_uncompress(b"\x05\x01\x00\x09")  # copy reaching before the start of the output


def test_entry_proto_bytes():
    """BundleEntryProto wire bytes assembled by hand: dtype=DT_FLOAT, shape [3,4], offset 12, size 48, crc 0x01020304."""
    got = tb._encode_entry(tb.DT_FLOAT, (3, 4), 12, 48, 0x01020304)
    want = bytes([0x08, 0x01,                                     # 1: dtype = 1
                  0x12, 0x08, 0x12, 0x02, 0x08, 0x03, 0x12, 0x02, 0x08, 0x04,   # 2: shape { dim{size:3} dim{size:4} }
                  0x20, 0x0C, 0x28, 0x30,                         # 4: offset = 12, 5: size = 48 (shard_id 0 omitted)
                  0x35, 0x04, 0x03, 0x02, 0x01])                  # 6: fixed32 crc32c
    assert got == want
    e = tb._decode_entry(want)
    assert (e["dtype"], e["shape"], e["offset"], e["size"], e["crc32c"], e["shard_id"]) == (1, (3, 4), 12, 48, 0x01020304, 0)
    assert tb._encode_header(1) == bytes([0x08, 0x01, 0x1A, 0x02, 0x08, 0x01])
    assert tb._decode_header(tb._encode_header(1)) == {"num_shards": 1, "endianness": 0, "producer": 1}
    scalar = tb._decode_entry(tb._encode_entry(tb.DT_INT64, (), 0, 8, 5))
    assert scalar["shape"] == () and scalar["dtype"] == tb.DT_INT64


def test_table_layout_small(tmp_path):
    """One data block, prefix-compressed keys, footer with the magic; bytes checked against the format by hand."""
    p = str(tmp_path / "t.index")
    tb.write_table(p, [(b"", b"H"), (b"ab", b"1"), (b"abc", b"22")])
    raw = open(p, "rb").read()
    block = (bytes([0, 0, 1]) + b"H" + bytes([0, 2, 1]) + b"ab" + b"1" + bytes([2, 1, 2]) + b"c" + b"22"
             + struct.pack("<II", 0, 1))
    assert raw[:len(block)] == block
    assert raw[len(block)] == 0  # kNoCompression
    crc = struct.unpack("<I", raw[len(block) + 1:len(block) + 5])[0]
    assert tb.unmask_crc(crc) == tb._crc32c_py(block + b"\x00")
    assert struct.unpack("<Q", raw[-8:])[0] == 0xDB4775248B80FB57 and len(raw[-48:]) == 48
    assert tb.read_table(p) == [(b"", b"H"), (b"ab", b"1"), (b"abc", b"22")]
    with pytest.raises(tb.BundleError):
        tb.write_table(p, [(b"b", b""), (b"a", b"")])
    # a flipped payload byte is caught by the block checksum
    bad = bytearray(raw)
    bad[4] ^= 1
    open(p, "wb").write(bad)
    with pytest.raises(tb.BundleError):
        tb.read_table(p)


def test_table_many_blocks_and_restarts(tmp_path):
    keys = sorted({("encoder/gmm_ll_%d/dense%s/%s" % (k, d, v)).encode() for k in range(90) for d in ("", "_1")
                   for v in ("kernel", "bias")})
    items = [(b"", b"hdr")] + [(k, k[::-1] * 3) for k in keys]
    p = str(tmp_path / "m.index")
    tb.write_table(p, items, block_size=512)  # forces many data blocks + index separators
    assert tb.read_table(p) == items
    assert tb._shortest_separator(b"abcdefg", b"abzz") == b"abd"
    assert tb._shortest_separator(b"abc", b"abd") == b"abc"
    assert tb._short_successor(b"\xff\xffa") == b"\xff\xffb"


def test_reads_snappy_compressed_index_blocks(tmp_path):
    """TF's TableBuilder may compress blocks (type 1): re-pack a written index with pyarrow's snappy and read it back."""
    pa = pytest.importorskip("pyarrow")
    items = [(b"", tb._encode_header(1))] + [(("v%03d/kernel" % i).encode(), tb._encode_entry(1, (2, 2), 16 * i, 16, i)) for i in range(40)]
    src = str(tmp_path / "a.index")
    tb.write_table(src, items)
    raw = open(src, "rb").read()
    footer = raw[-48:]
    pos = 0
    moff, pos = tb._get_varint(footer, pos)
    msize, pos = tb._get_varint(footer, pos)
    ioff, pos = tb._get_varint(footer, pos)
    isize, pos = tb._get_varint(footer, pos)
    (_, handle), = list(tb._block_entries(raw[ioff:ioff + isize]))
    doff, p2 = tb._get_varint(handle, 0)
    dsize, _ = tb._get_varint(handle, p2)

    out = bytearray()

    def emit(contents, typ):
        off = len(out)
        out.extend(contents)
        out.append(typ)
        out.extend(struct.pack("<I", tb.mask_crc(tb._crc32c_py(bytes([typ]), tb._crc32c_py(contents)))))
        return off, len(contents)

    d = emit(pa.compress(raw[doff:doff + dsize], codec="snappy", asbytes=True), 1)
    m = emit(raw[moff:moff + msize], 0)
    ib = tb._BlockBuilder(1)
    ib.add(b"w", tb._handle(*d))
    i = emit(ib.finish(), 0)
    f = tb._handle(*m) + tb._handle(*i)
    out.extend(f + b"\x00" * (40 - len(f)) + struct.pack("<Q", tb.TABLE_MAGIC))
    dst = str(tmp_path / "b.index")
    open(dst, "wb").write(out)
    assert tb.read_table(dst) == items


def test_bundle_roundtrip_reference_variable_names(tmp_path):
    rng = np.random.Generator(np.random.PCG64(9))
    state = {"imf_emb/kernel": rng.standard_normal((40, 8)).astype(np.float32),
             "imf_emb/bias": np.zeros(8, np.float32),
             "encoder/enc_embeddings": rng.standard_normal((30, 8)).astype(np.float32),
             "decoder/net/multi_rnn_cell/cell_0/lstm_cell/kernel": rng.standard_normal((24, 64)).astype(np.float32),
             "cnn/conv5_1/weights_conv": rng.standard_normal((3, 3, 4, 4)).astype(np.float32),
             "global_step": np.asarray(7, np.int64)}
    prefix = str(tmp_path / "checkpoints" / "default.ckpt")
    tb.write_bundle(prefix, state)
    assert os.path.exists(prefix + ".index") and os.path.exists(prefix + ".data-00000-of-00001")
    header, entries = tb.list_bundle(prefix)
    assert header == {"num_shards": 1, "endianness": 0, "producer": 1}
    assert list(entries) == sorted(state)  # key order = byte order of the names
    off = 0
    for n in sorted(state):  # tensors lie back to back in key order
        assert entries[n]["offset"] == off and entries[n]["shape"] == state[n].shape
        off += state[n].nbytes
    assert os.path.getsize(prefix + ".data-00000-of-00001") == off
    back = tb.read_bundle(prefix)
    for n in state:
        assert back[n].dtype == state[n].dtype and back[n].shape == state[n].shape
        np.testing.assert_array_equal(back[n], state[n])
    only = tb.read_bundle(prefix, ["imf_emb/bias"])
    assert list(only) == ["imf_emb/bias"]
    with pytest.raises(KeyError):
        tb.read_bundle(prefix, ["nope"])
    # corruption of the payload is caught by the per-tensor crc32c
    with open(prefix + ".data-00000-of-00001", "r+b") as f:
        f.seek(5)
        b = f.read(1)
        f.seek(5)
        f.write(bytes([b[0] ^ 0x40]))
    with pytest.raises(tb.BundleError):
        tb.read_bundle(prefix)
    tb.read_bundle(prefix, verify=False)
    tb.update_checkpoint_state(str(tmp_path / "checkpoints"), prefix)
    assert tb.latest_checkpoint(str(tmp_path / "checkpoints")) == prefix
    # Saver writes the path relative to the state file's directory (a moved checkpoint directory keeps working)
    assert open(str(tmp_path / "checkpoints" / "checkpoint")).read().startswith(
        'model_checkpoint_path: "%s"\n' % os.path.basename(prefix))
    with pytest.raises(FileNotFoundError):
        tb.list_bundle(str(tmp_path / "missing.ckpt"))
