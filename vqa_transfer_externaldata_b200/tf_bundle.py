"""TensorFlow checkpoint bundles (`model-N.index` + `model-N.data-00000-of-00001`) without TensorFlow.

The reference saves / restores its variables with tf.train.Saver (vqa/trainer.py:141-147, 173-186), i.e. in the
tensor-bundle format (tensorflow/core/util/tensor_bundle):
  * `<prefix>.index`: a LevelDB-style sorted string table (tensorflow/core/lib/io/table*): data blocks of
    prefix-compressed (key, value) entries with a restart array, each block followed by a 5-byte trailer
    (compression type 0 + masked CRC-32C of block + type), a metaindex block, an index block mapping the last key of
    every data block to its (offset, size) handle, and a 48-byte footer (two handles, padding, magic
    0xdb4775248b80fb57 little endian). Key "" holds BundleHeaderProto {1: num_shards, 2: endianness, 3: version};
    every other key is a variable name whose value is BundleEntryProto {1: dtype, 2: TensorShapeProto, 3: shard_id,
    4: offset, 5: size, 6: fixed32 masked crc32c}.
  * `<prefix>.data-SSSSS-of-NNNNN`: the tensors' raw little-endian bytes at [offset, offset + size).
This module reads and writes that format for the dtypes the path uses (float32, int32, int64) so that
Model.state_dict() / load_state_dict() can be filled from / saved as reference checkpoints by variable name (a pure
rename: parameters are kept in TF layout). The format is restated from its published definition; no TensorFlow
build was available to cross-check files, so tests pin the pieces that have known answers (CRC masking, footer magic,
prefix compression, proto encodings) and the writer / reader against each other.
"""
import os
import struct

import numpy as np

from .input_ops import _enc_varint, _fields, _varint, crc32c

TABLE_MAGIC = 0xDB4775248B80FB57
_MASK_DELTA = 0xA282EAD8
BLOCK_RESTART_INTERVAL = 16
BLOCK_SIZE = 4096
# tensorflow/core/framework/types.proto
DT_FLOAT, DT_INT32, DT_INT64 = 1, 3, 9
_NP = {DT_FLOAT: np.dtype("<f4"), DT_INT32: np.dtype("<i4"), DT_INT64: np.dtype("<i8")}
_DT = {np.dtype("float32"): DT_FLOAT, np.dtype("int32"): DT_INT32, np.dtype("int64"): DT_INT64}


def _mask(c):
    return (((c >> 15) | (c << 17)) + _MASK_DELTA) & 0xFFFFFFFF


# ---------------------------------------------------------------------------------------------------------------------
# sorted string table
# ---------------------------------------------------------------------------------------------------------------------
def _parse_block(buf):
    """Entries of one block: shared-prefix varint, non-shared varint, value-length varint, key delta, value; the last
    4 bytes hold the number of restarts, preceded by that many fixed32 restart offsets."""
    n_restarts = struct.unpack("<I", buf[-4:])[0]
    end = len(buf) - 4 - 4 * n_restarts
    pos, key, out = 0, b"", []
    while pos < end:
        shared, pos = _varint(buf, pos)
        non_shared, pos = _varint(buf, pos)
        vlen, pos = _varint(buf, pos)
        key = key[:shared] + bytes(buf[pos:pos + non_shared])
        pos += non_shared
        out.append((key, bytes(buf[pos:pos + vlen])))
        pos += vlen
    return out


def _read_block(data, offset, size, verify=True):
    block = data[offset:offset + size]
    trailer = data[offset + size:offset + size + 5]
    if len(block) < size or len(trailer) < 5:
        raise ValueError("truncated table block")
    if trailer[0] != 0:
        raise ValueError("compressed table blocks are not supported (tf.train.Saver writes none)")
    if verify and _mask(crc32c(block + trailer[:1])) != struct.unpack("<I", trailer[1:])[0]:
        raise ValueError("table block checksum mismatch")
    return _parse_block(block)


def _handle(buf, pos):
    off, pos = _varint(buf, pos)
    size, pos = _varint(buf, pos)
    return off, size, pos


def read_table(path, verify=True):
    """All (key, value) pairs of a table file, in key order."""
    with open(path, "rb") as f:
        data = f.read()
    if len(data) < 48:
        raise ValueError(f"{path}: too short for a table footer")
    footer = data[-48:]
    if struct.unpack("<Q", footer[40:])[0] != TABLE_MAGIC:
        raise ValueError(f"{path}: bad table magic")
    _, _, pos = _handle(footer, 0)              # metaindex handle (unused)
    ioff, isize, _ = _handle(footer, pos)       # index handle
    out = []
    for _, handle in _read_block(data, ioff, isize, verify):
        off, size, _ = _handle(handle, 0)
        out.extend(_read_block(data, off, size, verify))
    return out


class _BlockBuilder:
    def __init__(self):
        self.buf, self.restarts, self.count, self.last = bytearray(), [0], 0, b""

    def add(self, key, value):
        shared = 0
        if self.count % BLOCK_RESTART_INTERVAL == 0 and self.count:
            self.restarts.append(len(self.buf))
        elif self.count:
            n = min(len(key), len(self.last))
            while shared < n and key[shared] == self.last[shared]:
                shared += 1
        self.buf += _enc_varint(shared) + _enc_varint(len(key) - shared) + _enc_varint(len(value))
        self.buf += key[shared:] + value
        self.last, self.count = key, self.count + 1

    def finish(self):
        return bytes(self.buf) + b"".join(struct.pack("<I", r) for r in self.restarts) + struct.pack("<I", len(self.restarts))


def write_table(path, items):
    """items: (key bytes, value bytes) pairs; written in sorted key order."""
    out = bytearray()
    index = _BlockBuilder()

    def emit(block_bytes):
        off = len(out)
        out.extend(block_bytes)
        out.extend(b"\x00" + struct.pack("<I", _mask(crc32c(block_bytes + b"\x00"))))
        return off, len(block_bytes)

    cur = _BlockBuilder()
    for key, value in sorted(items):
        cur.add(key, value)
        if len(cur.buf) >= BLOCK_SIZE:
            off, size = emit(cur.finish())
            index.add(cur.last, _enc_varint(off) + _enc_varint(size))
            cur = _BlockBuilder()
    if cur.count:
        off, size = emit(cur.finish())
        index.add(cur.last, _enc_varint(off) + _enc_varint(size))
    moff, msize = emit(_BlockBuilder().finish())            # empty metaindex block
    ioff, isize = emit(index.finish())
    footer = _enc_varint(moff) + _enc_varint(msize) + _enc_varint(ioff) + _enc_varint(isize)
    footer += b"\x00" * (40 - len(footer)) + struct.pack("<Q", TABLE_MAGIC)
    out.extend(footer)
    with open(path, "wb") as f:
        f.write(bytes(out))


# ---------------------------------------------------------------------------------------------------------------------
# bundle protos
# ---------------------------------------------------------------------------------------------------------------------
def _enc_field(num, wt, payload):
    return _enc_varint((num << 3) | wt) + payload


def _encode_entry(dtype, shape, offset, size, crc):
    dims = b"".join(_enc_field(2, 2, _enc_varint(len(d)) + d) for d in (_enc_field(1, 0, _enc_varint(int(n))) for n in shape))
    e = _enc_field(1, 0, _enc_varint(dtype)) + _enc_field(2, 2, _enc_varint(len(dims)) + dims)
    if offset:
        e += _enc_field(4, 0, _enc_varint(offset))       # shard_id 0 and offset 0 are proto3 defaults: omitted
    e += _enc_field(5, 0, _enc_varint(size)) + _enc_field(6, 5, struct.pack("<I", crc))
    return e


def _decode_entry(buf):
    ent = {"dtype": 0, "shape": [], "shard_id": 0, "offset": 0, "size": 0, "crc32c": None}
    for num, wt, v in _fields(memoryview(buf)):
        if num == 1 and wt == 0:
            ent["dtype"] = v
        elif num == 2 and wt == 2:
            for n2, w2, dim in _fields(v):
                if n2 == 2 and w2 == 2:
                    size = 0
                    for n3, w3, x in _fields(dim):
                        if n3 == 1 and w3 == 0:
                            size = x
                    ent["shape"].append(size)
        elif num == 3 and wt == 0:
            ent["shard_id"] = v
        elif num == 4 and wt == 0:
            ent["offset"] = v
        elif num == 5 and wt == 0:
            ent["size"] = v
        elif num == 6 and wt == 5:
            ent["crc32c"] = struct.unpack("<I", bytes(v))[0]
    return ent


def read_bundle(prefix, names=None, verify=True):
    """dict variable name -> ndarray for the variables of the bundle `prefix` (e.g. '.../model-12000')."""
    entries = read_table(prefix + ".index", verify)
    header = {"num_shards": 1}
    out = {}
    shards = {}
    for key, value in entries:
        if key == b"":
            for num, wt, v in _fields(memoryview(value)):
                if num == 1 and wt == 0:
                    header["num_shards"] = v
                elif num == 2 and wt == 0 and v != 0:
                    raise ValueError("big-endian bundles are not supported")
            continue
        name = key.decode("utf-8")
        if names is not None and name not in names:
            continue
        ent = _decode_entry(value)
        if ent["dtype"] not in _NP:
            if names is None:
                continue                        # e.g. string tensors of a Saver's bookkeeping: not part of the path
            raise ValueError(f"{name}: dtype {ent['dtype']} is not supported")
        sid = ent["shard_id"]
        if sid not in shards:
            p = f"{prefix}.data-{sid:05d}-of-{header['num_shards']:05d}"
            with open(p, "rb") as f:
                shards[sid] = f.read()
        raw = shards[sid][ent["offset"]:ent["offset"] + ent["size"]]
        if len(raw) != ent["size"]:
            raise ValueError(f"{name}: data shard is truncated")
        if verify and ent["crc32c"] is not None and _mask(crc32c(raw)) != ent["crc32c"]:
            raise ValueError(f"{name}: tensor checksum mismatch")
        a = np.frombuffer(raw, dtype=_NP[ent["dtype"]])
        if int(np.prod(ent["shape"], dtype=np.int64)) != a.size:
            raise ValueError(f"{name}: shape {ent['shape']} does not match {a.size} elements")
        out[name] = a.reshape(ent["shape"]).copy()
    return out


def write_bundle(prefix, tensors):
    """Write dict variable name -> array as a one-shard bundle."""
    os.makedirs(os.path.dirname(os.path.abspath(prefix)), exist_ok=True)
    items = [(b"", _enc_field(1, 0, _enc_varint(1)) + _enc_field(3, 2, _enc_varint(2) + _enc_field(1, 0, _enc_varint(1))))]
    data = bytearray()
    for name in sorted(tensors):
        a = np.asarray(tensors[name])
        if a.ndim and not a.flags.c_contiguous:     # (np.ascontiguousarray would turn a scalar into shape [1])
            a = np.ascontiguousarray(a)
        if a.dtype not in _DT:
            raise ValueError(f"{name}: dtype {a.dtype} is not supported")
        raw = a.astype(_NP[_DT[a.dtype]], copy=False).tobytes()
        items.append((name.encode("utf-8"), _encode_entry(_DT[a.dtype], a.shape, len(data), len(raw), _mask(crc32c(raw)))))
        data += raw
    write_table(prefix + ".index", items)
    with open(prefix + ".data-00000-of-00001", "wb") as f:
        f.write(bytes(data))
