"""Reader and writer for TensorFlow's V2 checkpoint format ("tensor bundle"), written without TensorFlow.

This is what `tf.train.Saver.save / restore` puts behind `./checkpoints/{name}.ckpt` (reference: main.py:186-191, 211,
288; ops/inference.py:6-7; gen_caption.py:113-115), so checkpoints trained by the reference load into this path and
checkpoints written here restore in the reference by variable name.

On-disk layout (tensorflow/core/util/tensor_bundle, tensorflow/core/lib/io/{table,block,format}: the published
LevelDB table format with TF's options):

  {prefix}.data-00000-of-00001   raw little-endian tensor bytes, back to back, in key order
  {prefix}.index                 an SSTable: key "" -> BundleHeaderProto, key <variable name> -> BundleEntryProto
      data blocks   : entries `varint shared | varint non_shared | varint value_len | key suffix | value`,
                      a restart point (shared = 0) every 16 entries, then the uint32 restart offsets + their count;
                      each block is followed by a 5-byte trailer: compression type (0 raw, 1 snappy) and the masked
                      crc32c of block + type
      metaindex     : an empty block;  index block: separator key -> BlockHandle(varint offset, varint size)
      footer        : the two handles padded to 40 bytes + the magic 0xdb4775248b80fb57 (little endian), 48 bytes

  BundleHeaderProto { int32 num_shards = 1; enum endianness = 2; VersionDef version = 3 { int32 producer = 1; } }
  BundleEntryProto  { enum dtype = 1; TensorShapeProto shape = 2 { repeated Dim dim = 2 { int64 size = 1; } }
                      int32 shard_id = 3; int64 offset = 4; int64 size = 5; fixed32 crc32c = 6 (masked);
                      repeated TensorSliceProto slices = 7; }

TF itself may snappy-compress index blocks, so the reader carries a snappy decompressor; the writer stores raw blocks
(every TF reader accepts both). The crc32c of the tensor payloads (hundreds of MB with the cnn/ variables) is computed by
the host helper `vc_crc32c` of libvaecap.so; a table-driven pure-Python version is used only when the library is absent.
"""
import os
import struct

import numpy as np

TABLE_MAGIC = 0xDB4775248B80FB57
_MASK_DELTA = 0xA282EAD8
BLOCK_SIZE = 262144        # table::Options::block_size in TF
RESTART_INTERVAL = 16      # table::Options::block_restart_interval

# tensorflow/core/framework/types.proto
DT_FLOAT, DT_DOUBLE, DT_INT32, DT_UINT8, DT_INT64, DT_BOOL, DT_BFLOAT16, DT_HALF = 1, 2, 3, 4, 9, 10, 14, 19
_NP_OF_DT = {DT_FLOAT: np.dtype("<f4"), DT_DOUBLE: np.dtype("<f8"), DT_INT32: np.dtype("<i4"), DT_UINT8: np.dtype("u1"),
             DT_INT64: np.dtype("<i8"), DT_BOOL: np.dtype("?"), DT_HALF: np.dtype("<f2"), DT_BFLOAT16: np.dtype("<u2")}
_DT_OF_NP = {v: k for k, v in _NP_OF_DT.items() if k != DT_BFLOAT16}


class BundleError(ValueError):
    pass


# ----------------------------------------------------------------------------------------------- crc32c (Castagnoli)
_CRC_TABLE = None


def _crc32c_py(data, crc=0):
    global _CRC_TABLE
    if _CRC_TABLE is None:
        tab = []
        for i in range(256):
            c = i
            for _ in range(8):
                c = (c >> 1) ^ 0x82F63B78 if c & 1 else c >> 1
            tab.append(c)
        _CRC_TABLE = tab
    tab = _CRC_TABLE
    c = crc ^ 0xFFFFFFFF
    for b in bytes(data):
        c = tab[(c ^ b) & 0xFF] ^ (c >> 8)
    return c ^ 0xFFFFFFFF


def crc32c(data, crc=0):
    """crc32c of a bytes-like / contiguous ndarray, continuing from `crc`."""
    buf = np.frombuffer(data, dtype=np.uint8) if not isinstance(data, np.ndarray) else data.reshape(-1).view(np.uint8)
    if buf.size < 4096:
        return _crc32c_py(buf.tobytes(), crc)
    try:
        import ctypes
        from . import lib as L
        fn = L.load().vc_crc32c
        fn.restype = ctypes.c_uint32
        fn.argtypes = [ctypes.c_uint32, ctypes.c_void_p, ctypes.c_size_t]
        buf = np.ascontiguousarray(buf)
        return int(fn(crc, buf.ctypes.data, buf.size))
    except (OSError, AttributeError, RuntimeError):
        return _crc32c_py(buf.tobytes(), crc)


def mask_crc(c):
    return ((((c >> 15) | (c << 17)) & 0xFFFFFFFF) + _MASK_DELTA) & 0xFFFFFFFF


def unmask_crc(m):
    r = (m - _MASK_DELTA) & 0xFFFFFFFF
    return ((r >> 17) | (r << 15)) & 0xFFFFFFFF


# ----------------------------------------------------------------------------------------------- varints / protobuf
def _put_varint(out, v):
    v &= (1 << 64) - 1
    while v >= 0x80:
        out.append((v & 0x7F) | 0x80)
        v >>= 7
    out.append(v)


def _get_varint(buf, pos):
    shift = v = 0
    while True:
        if pos >= len(buf):
            raise BundleError("truncated varint")
        b = buf[pos]
        pos += 1
        v |= (b & 0x7F) << shift
        if b < 0x80:
            return v, pos
        shift += 7
        if shift > 63:
            raise BundleError("varint too long")


def _pb_fields(buf):
    """Yields (field number, wire type, value) of a serialized message; length-delimited values as bytes."""
    pos = 0
    buf = bytes(buf)
    while pos < len(buf):
        key, pos = _get_varint(buf, pos)
        fno, wt = key >> 3, key & 7
        if wt == 0:
            v, pos = _get_varint(buf, pos)
        elif wt == 1:
            v = struct.unpack_from("<Q", buf, pos)[0]
            pos += 8
        elif wt == 2:
            n, pos = _get_varint(buf, pos)
            v = buf[pos:pos + n]
            if len(v) != n:
                raise BundleError("truncated protobuf field")
            pos += n
        elif wt == 5:
            v = struct.unpack_from("<I", buf, pos)[0]
            pos += 4
        else:
            raise BundleError("unsupported protobuf wire type %d" % wt)
        yield fno, wt, v


def _pb_varint_field(out, fno, v):
    if v:  # proto3: zero scalars are not serialised
        _put_varint(out, fno << 3)
        _put_varint(out, v)


def _pb_bytes_field(out, fno, payload):
    _put_varint(out, (fno << 3) | 2)
    _put_varint(out, len(payload))
    out.extend(payload)


def _encode_header(num_shards=1):
    out = bytearray()
    _pb_varint_field(out, 1, num_shards)          # endianness LITTLE = 0 is the default and is omitted
    ver = bytearray()
    _pb_varint_field(ver, 1, 1)                   # VersionDef.producer = kTensorBundleVersion
    _pb_bytes_field(out, 3, ver)
    return bytes(out)


def _encode_entry(dt, shape, offset, size, crc_masked, shard_id=0):
    out = bytearray()
    _pb_varint_field(out, 1, dt)
    shp = bytearray()
    for d in shape:
        dim = bytearray()
        _pb_varint_field(dim, 1, int(d))
        _pb_bytes_field(shp, 2, dim)
    _pb_bytes_field(out, 2, shp)
    _pb_varint_field(out, 3, shard_id)
    _pb_varint_field(out, 4, offset)
    _pb_varint_field(out, 5, size)
    _put_varint(out, (6 << 3) | 5)
    out.extend(struct.pack("<I", crc_masked))
    return bytes(out)


def _signed64(v):
    return v - (1 << 64) if v >= (1 << 63) else v


def _decode_entry(buf):
    e = {"dtype": 0, "shape": (), "shard_id": 0, "offset": 0, "size": 0, "crc32c": None, "slices": 0}
    for fno, _, v in _pb_fields(buf):
        if fno == 1:
            e["dtype"] = v
        elif fno == 2:
            dims = []
            for f2, _, v2 in _pb_fields(v):
                if f2 == 2:
                    size = 0
                    for f3, _, v3 in _pb_fields(v2):
                        if f3 == 1:
                            size = _signed64(v3)
                    dims.append(size)
                elif f2 == 3 and v2:
                    raise BundleError("tensor of unknown rank in checkpoint")
            e["shape"] = tuple(dims)
        elif fno == 3:
            e["shard_id"] = v
        elif fno == 4:
            e["offset"] = v
        elif fno == 5:
            e["size"] = v
        elif fno == 6:
            e["crc32c"] = v
        elif fno == 7:
            e["slices"] += 1
    return e


def _decode_header(buf):
    h = {"num_shards": 0, "endianness": 0, "producer": 0}
    for fno, _, v in _pb_fields(buf):
        if fno == 1:
            h["num_shards"] = v
        elif fno == 2:
            h["endianness"] = v
        elif fno == 3:
            for f2, _, v2 in _pb_fields(v):
                if f2 == 1:
                    h["producer"] = v2
    return h


# ----------------------------------------------------------------------------------------------- snappy (read side)
def snappy_uncompress(src):
    """Raw snappy block format: varint uncompressed length, then literal / copy elements."""
    src = bytes(src)
    n, pos = _get_varint(src, 0)
    out = bytearray()
    while pos < len(src):
        tag = src[pos]
        pos += 1
        kind = tag & 3
        if kind == 0:
            ln = tag >> 2
            if ln >= 60:
                nb = ln - 59
                ln = int.from_bytes(src[pos:pos + nb], "little")
                pos += nb
            ln += 1
            out += src[pos:pos + ln]
            pos += ln
            continue
        if kind == 1:
            ln = ((tag >> 2) & 7) + 4
            off = ((tag >> 5) << 8) | src[pos]
            pos += 1
        elif kind == 2:
            ln = (tag >> 2) + 1
            off = src[pos] | (src[pos + 1] << 8)
            pos += 2
        else:
            ln = (tag >> 2) + 1
            off = int.from_bytes(src[pos:pos + 4], "little")
            pos += 4
        if off == 0 or off > len(out):
            raise BundleError("corrupt snappy stream")
        start = len(out) - off
        for i in range(ln):  # copies may overlap their own output
            out.append(out[start + i])
    if len(out) != n:
        raise BundleError("snappy length mismatch: header %d, got %d" % (n, len(out)))
    return bytes(out)


# ----------------------------------------------------------------------------------------------- SSTable
class _BlockBuilder(object):
    def __init__(self, restart_interval):
        self.ri = restart_interval
        self.reset()

    def reset(self):
        self.buf = bytearray()
        self.restarts = [0]
        self.count = 0
        self.last = b""

    def add(self, key, value):
        shared = 0
        if self.count < self.ri:
            m = min(len(key), len(self.last))
            while shared < m and key[shared] == self.last[shared]:
                shared += 1
        else:
            self.restarts.append(len(self.buf))
            self.count = 0
        _put_varint(self.buf, shared)
        _put_varint(self.buf, len(key) - shared)
        _put_varint(self.buf, len(value))
        self.buf += key[shared:]
        self.buf += value
        self.last = key
        self.count += 1

    def size_estimate(self):
        return len(self.buf) + 4 * len(self.restarts) + 4

    def empty(self):
        return not self.buf

    def finish(self):
        return bytes(self.buf) + b"".join(struct.pack("<I", r) for r in self.restarts) + struct.pack("<I", len(self.restarts))


def _shortest_separator(start, limit):
    """BytewiseComparator::FindShortestSeparator: a short key k with start <= k < limit."""
    m = min(len(start), len(limit))
    d = 0
    while d < m and start[d] == limit[d]:
        d += 1
    if d < m and start[d] < 0xFF and start[d] + 1 < limit[d]:
        return start[:d] + bytes([start[d] + 1])
    return start


def _short_successor(key):
    for i, b in enumerate(key):
        if b != 0xFF:
            return key[:i] + bytes([b + 1])
    return key


def _handle(offset, size):
    out = bytearray()
    _put_varint(out, offset)
    _put_varint(out, size)
    return bytes(out)


def write_table(path, items, block_size=BLOCK_SIZE):
    """items: iterable of (key bytes, value bytes) in strictly increasing bytewise key order."""
    out = bytearray()
    data, index = _BlockBuilder(RESTART_INTERVAL), _BlockBuilder(1)

    def emit(block):
        contents = block.finish()
        off = len(out)
        out.extend(contents)
        out.append(0)  # kNoCompression
        out.extend(struct.pack("<I", mask_crc(_crc32c_py(b"\x00", _crc32c_py(contents)))))
        block.reset()
        return off, len(contents)

    pending = None  # (last key of the finished block, handle): its index key needs the next block's first key
    last_key = None
    for key, value in items:
        if last_key is not None and not key > last_key:
            raise BundleError("table keys must be strictly increasing: %r after %r" % (key, last_key))
        if pending is not None:
            index.add(_shortest_separator(pending[0], key), pending[1])
            pending = None
        data.add(key, value)
        last_key = key
        if data.size_estimate() >= block_size:
            pending = (last_key, _handle(*emit(data)))
    if not data.empty():
        pending = (last_key, _handle(*emit(data)))
    if pending is not None:
        index.add(_short_successor(pending[0]), pending[1])
    meta_handle = _handle(*emit(_BlockBuilder(RESTART_INTERVAL)))
    index_handle = _handle(*emit(index))
    footer = meta_handle + index_handle
    footer += b"\x00" * (40 - len(footer)) + struct.pack("<Q", TABLE_MAGIC)
    out.extend(footer)
    with open(path, "wb") as f:
        f.write(out)


def _read_block(buf, offset, size, verify=True):
    raw = buf[offset:offset + size]
    trailer = buf[offset + size:offset + size + 5]
    if len(raw) != size or len(trailer) != 5:
        raise BundleError("table block out of range")
    if verify:
        want = unmask_crc(struct.unpack("<I", trailer[1:])[0])
        if _crc32c_py(trailer[:1], _crc32c_py(raw)) != want:
            raise BundleError("block checksum mismatch")
    if trailer[0] == 1:
        raw = snappy_uncompress(raw)
    elif trailer[0] != 0:
        raise BundleError("unknown block compression type %d" % trailer[0])
    return raw


def _block_entries(block):
    if len(block) < 4:
        raise BundleError("bad block")
    nrestarts = struct.unpack_from("<I", block, len(block) - 4)[0]
    end = len(block) - 4 - 4 * nrestarts
    if end < 0:
        raise BundleError("bad block restart array")
    pos, key = 0, b""
    while pos < end:
        shared, pos = _get_varint(block, pos)
        non_shared, pos = _get_varint(block, pos)
        vlen, pos = _get_varint(block, pos)
        if shared > len(key) or pos + non_shared + vlen > end:
            raise BundleError("corrupt block entry")
        key = key[:shared] + block[pos:pos + non_shared]
        pos += non_shared
        value = block[pos:pos + vlen]
        pos += vlen
        yield key, value


def read_table(path, verify=True):
    """-> list of (key, value) of an SSTable file, in key order."""
    with open(path, "rb") as f:
        buf = f.read()
    if len(buf) < 48:
        raise BundleError("%s: too short to be a table" % path)
    footer = buf[-48:]
    if struct.unpack("<Q", footer[40:])[0] != TABLE_MAGIC:
        raise BundleError("%s: bad table magic (not a TF V2 checkpoint index)" % path)
    pos = 0
    _, pos = _get_varint(footer, pos)
    _, pos = _get_varint(footer, pos)      # metaindex handle: unused by the bundle
    ioff, pos = _get_varint(footer, pos)
    isize, pos = _get_varint(footer, pos)
    items = []
    for _, handle in _block_entries(_read_block(buf, ioff, isize, verify)):
        off, p = _get_varint(handle, 0)
        size, p = _get_varint(handle, p)
        items.extend(_block_entries(_read_block(buf, off, size, verify)))
    return items


# ----------------------------------------------------------------------------------------------- the bundle
def data_path(prefix, shard=0, num_shards=1):
    return "%s.data-%05d-of-%05d" % (prefix, shard, num_shards)


def write_bundle(prefix, tensors):
    """tensors: {variable name: ndarray}. Writes {prefix}.index and {prefix}.data-00000-of-00001 (one shard)."""
    d = os.path.dirname(prefix)
    if d and not os.path.exists(d):
        os.makedirs(d)
    names = sorted(tensors, key=lambda s: s.encode("utf-8"))
    if "" in tensors:
        raise BundleError("the empty name is reserved for the bundle header")
    items = [(b"", _encode_header(1))]
    offset = 0
    tmp = data_path(prefix) + ".tempstate"
    with open(tmp, "wb") as f:
        for name in names:
            a = np.asarray(tensors[name])
            shape = a.shape  # ascontiguousarray would turn a scalar into shape (1,)
            le = a.dtype.newbyteorder("<") if a.dtype.byteorder == ">" else a.dtype
            a = np.ascontiguousarray(a, dtype=le)
            dt = _DT_OF_NP.get(np.dtype(a.dtype.str.replace("=", "<")) if a.dtype.itemsize > 1 else a.dtype)
            if dt is None:
                raise BundleError("variable %s: dtype %s has no checkpoint encoding here" % (name, a.dtype))
            raw = a.reshape(-1).view(np.uint8)
            f.write(raw.data)
            items.append((name.encode("utf-8"), _encode_entry(dt, shape, offset, raw.size, mask_crc(crc32c(raw)))))
            offset += raw.size
    os.replace(tmp, data_path(prefix))
    tmp = prefix + ".index.tempstate"
    write_table(tmp, items)
    os.replace(tmp, prefix + ".index")
    return prefix


def list_bundle(prefix, verify=True):
    """-> (header dict, {name: entry dict}) of {prefix}.index."""
    idx = prefix + ".index"
    if not os.path.exists(idx):
        raise FileNotFoundError(idx)
    items = read_table(idx, verify)
    if not items or items[0][0] != b"":
        raise BundleError("%s: no bundle header entry" % idx)
    header = _decode_header(items[0][1])
    if header["endianness"] != 0:
        raise BundleError("%s: big-endian bundles are not supported" % idx)
    return header, {k.decode("utf-8"): _decode_entry(v) for k, v in items[1:]}


def read_bundle(prefix, names=None, verify=True):
    """-> {name: ndarray} for `names` (default: every tensor). verify checks the block and tensor crc32c's."""
    header, entries = list_bundle(prefix, verify)
    want = list(entries) if names is None else list(names)
    out = {}
    files = {}
    try:
        for name in want:
            if name not in entries:
                raise KeyError("tensor %s not found in checkpoint %s" % (name, prefix))
            e = entries[name]
            if e["slices"]:
                raise NotImplementedError("%s is a partitioned variable; sliced entries are not supported" % name)
            if e["dtype"] not in _NP_OF_DT:
                raise BundleError("%s: unsupported dtype enum %d" % (name, e["dtype"]))
            dt = _NP_OF_DT[e["dtype"]]
            count = int(np.prod(e["shape"], dtype=np.int64)) if e["shape"] else 1
            if count * dt.itemsize != e["size"]:
                raise BundleError("%s: entry size %d does not match shape %s" % (name, e["size"], e["shape"]))
            sh = e["shard_id"]
            if sh not in files:
                files[sh] = open(data_path(prefix, sh, max(header["num_shards"], 1)), "rb")
            f = files[sh]
            f.seek(e["offset"])
            a = np.fromfile(f, dtype=dt, count=count)
            if a.size != count:
                raise BundleError("%s: data file truncated" % name)
            if verify and e["crc32c"] is not None and crc32c(a) != unmask_crc(e["crc32c"]):
                raise BundleError("%s: tensor checksum mismatch" % name)
            out[name] = a.reshape(e["shape"])
    finally:
        for f in files.values():
            f.close()
    return out


def update_checkpoint_state(directory, prefix, keep=()):
    """The text-format CheckpointState file `checkpoint` that Saver.save maintains next to the bundles."""
    # Saver writes paths RELATIVE to the state file's directory when the bundle lives there (so a checkpoint directory
    # can be moved); tf.train.latest_checkpoint joins them back
    def rel(p):
        return os.path.basename(p) if os.path.abspath(os.path.dirname(p) or ".") == os.path.abspath(directory) else p
    paths = [p for p in keep if p != prefix] + [prefix]
    with open(os.path.join(directory, "checkpoint"), "w") as f:
        f.write('model_checkpoint_path: "%s"\n' % rel(prefix))
        for p in paths:
            f.write('all_model_checkpoint_paths: "%s"\n' % rel(p))


def latest_checkpoint(directory):
    """tf.train.latest_checkpoint: the model_checkpoint_path of `directory/checkpoint`, or None."""
    p = os.path.join(directory, "checkpoint")
    if not os.path.exists(p):
        return None
    with open(p) as f:
        for line in f:
            if line.startswith("model_checkpoint_path:"):
                v = line.split(":", 1)[1].strip().strip('"')
                return v if os.path.isabs(v) else os.path.join(directory, v)
    return None
