"""Input side of the path: the batches vqa/datasets/input_ops_vqa_tf_record_memft.py:6-82 feeds the model,
read from the same TFRecord shards WITHOUT TensorFlow.

    create(batch_size, tf_record_dir, split, is_train=True, shuffle=True) -> iterator of batch dicts
        {'id' i64 [B], 'image_id' bytes [B], 'image_idx' i64 [B], 'q_intseq' i32 [B, T] (pad id 0, T = longest of the
         batch), 'q_intseq_len' i32 [B], 'answer_target' f32 [B, num_answers]}

What is restated here (published formats, no reference code involved):
  * TFRecord framing: u64 length | u32 masked crc32c(length) | payload | u32 masked crc32c(payload), little endian;
    mask(c) = ((c >> 15 | c << 17) + 0xa282ead8) mod 2^32, CRC-32C (Castagnoli, reflected polynomial 0x82F63B78);
  * tf.train.Example: Example{1: Features{1: map<string, Feature>}}, Feature = oneof {1: BytesList, 2: FloatList,
    3: Int64List}, each {1: repeated value} (packed or not) in protobuf wire format;
  * the parse_fn of the reference: FixedLenFeature defaults (qid -1, image_id "", image_idx -1),
    sparse_to_dense(answers/ids, [num_answers], answers/scores) for the soft-score target, padded_batch.
The records are written by data/tools/vqa_v2/generator_tf_record_memft_genome.py:184-195; `write_shards` below writes
the same layout (used by the tests and by users who have no TensorFlow to generate data with).
The feature bank itself is resident in HBM and indexed by `image_idx` on the device (vqa_forward's gather), so this
module only has to deliver ~12 KB per sample.
"""
import glob
import os
import struct

import numpy as np

# ---------------------------------------------------------------------------------------------------------------------
# CRC-32C
# ---------------------------------------------------------------------------------------------------------------------
_POLY = 0x82F63B78
_TABLE = []
for _i in range(256):
    _c = _i
    for _ in range(8):
        _c = (_c >> 1) ^ _POLY if _c & 1 else _c >> 1
    _TABLE.append(_c)
_MASK_DELTA = 0xA282EAD8


def crc32c(data):
    """CRC-32C of a bytes-like object (pure Python table walk; the C library's vqa_crc32c is used when it is loaded)."""
    fast = _fast_crc()
    if fast is not None:
        return fast(bytes(data))
    c = 0xFFFFFFFF
    for b in bytes(data):
        c = _TABLE[(c ^ b) & 0xFF] ^ (c >> 8)
    return c ^ 0xFFFFFFFF


_fast = False


def _fast_crc():
    global _fast
    if _fast is False:
        _fast = None
        try:
            import ctypes as C
            from . import lib as L
            if os.path.exists(L.LIB_PATH):
                fn = L.load().vqa_crc32c
                _fast = lambda b: int(fn(b, C.c_uint64(len(b))))  # noqa: E731
        except Exception:  # noqa: BLE001 -- the table walk above is the same function
            _fast = None
    return _fast


def masked_crc(data):
    c = crc32c(data)
    return (((c >> 15) | (c << 17)) + _MASK_DELTA) & 0xFFFFFFFF


# ---------------------------------------------------------------------------------------------------------------------
# TFRecord framing
# ---------------------------------------------------------------------------------------------------------------------
def read_records(path, verify_crc=True):
    """Yield the payload of every record of one TFRecord file."""
    with open(path, "rb") as f:
        while True:
            head = f.read(12)
            if not head:
                return
            if len(head) < 12:
                raise ValueError(f"{path}: truncated record header")
            (length,), (lcrc,) = struct.unpack("<Q", head[:8]), struct.unpack("<I", head[8:])
            if verify_crc and masked_crc(head[:8]) != lcrc:
                raise ValueError(f"{path}: corrupted record length")
            body = f.read(length + 4)
            if len(body) < length + 4:
                raise ValueError(f"{path}: truncated record")
            data = body[:length]
            if verify_crc and masked_crc(data) != struct.unpack("<I", body[length:])[0]:
                raise ValueError(f"{path}: corrupted record payload")
            yield data


def write_records(path, payloads):
    with open(path, "wb") as f:
        for data in payloads:
            head = struct.pack("<Q", len(data))
            f.write(head + struct.pack("<I", masked_crc(head)) + data + struct.pack("<I", masked_crc(data)))


# ---------------------------------------------------------------------------------------------------------------------
# protobuf wire format (only what tf.train.Example needs)
# ---------------------------------------------------------------------------------------------------------------------
def _varint(buf, pos):
    result, shift = 0, 0
    while True:
        b = buf[pos]
        pos += 1
        result |= (b & 0x7F) << shift
        if not b & 0x80:
            return result, pos
        shift += 7
        if shift > 70:
            raise ValueError("malformed varint")


def _fields(buf):
    """Yield (field number, wire type, value) of one message; value = int (varint, fixed) or a memoryview slice."""
    pos, n = 0, len(buf)
    while pos < n:
        key, pos = _varint(buf, pos)
        num, wt = key >> 3, key & 7
        if wt == 0:
            v, pos = _varint(buf, pos)
        elif wt == 1:
            v, pos = buf[pos:pos + 8], pos + 8
        elif wt == 2:
            ln, pos = _varint(buf, pos)
            v, pos = buf[pos:pos + ln], pos + ln
        elif wt == 5:
            v, pos = buf[pos:pos + 4], pos + 4
        else:
            raise ValueError(f"unsupported wire type {wt}")
        if pos > n:
            raise ValueError("truncated message")
        yield num, wt, v


def _signed64(v):
    return v - (1 << 64) if v >= (1 << 63) else v


def _parse_feature(buf):
    """Feature -> ('bytes', [bytes]) | ('float', float32 array) | ('int64', int64 array)."""
    for num, wt, v in _fields(buf):
        if wt != 2:
            continue
        if num == 1:
            return "bytes", [bytes(x) for n2, w2, x in _fields(v) if n2 == 1 and w2 == 2]
        if num == 2:
            vals = []
            for n2, w2, x in _fields(v):
                if n2 != 1:
                    continue
                if w2 == 2:      # packed
                    vals.append(np.frombuffer(x, dtype="<f4"))
                elif w2 == 5:    # one fixed32 per element
                    vals.append(np.frombuffer(x, dtype="<f4"))
            return "float", (np.concatenate(vals) if vals else np.zeros(0, np.float32)).astype(np.float32)
        if num == 3:
            vals = []
            for n2, w2, x in _fields(v):
                if n2 != 1:
                    continue
                if w2 == 2:
                    p = 0
                    while p < len(x):
                        y, p = _varint(x, p)
                        vals.append(_signed64(y))
                elif w2 == 0:
                    vals.append(_signed64(x))
            return "int64", np.asarray(vals, dtype=np.int64)
    return "none", None


def parse_example(data):
    """Serialized tf.train.Example -> dict feature name -> (kind, value)."""
    out = {}
    buf = memoryview(data)
    for num, wt, feats in _fields(buf):
        if num != 1 or wt != 2:
            continue
        for n2, w2, entry in _fields(feats):
            if n2 != 1 or w2 != 2:
                continue
            key, val = None, None
            for n3, w3, x in _fields(entry):
                if n3 == 1 and w3 == 2:
                    key = bytes(x).decode("utf-8")
                elif n3 == 2 and w3 == 2:
                    val = _parse_feature(x)
            if key is not None:
                out[key] = val if val is not None else ("none", None)
    return out


def _enc_varint(v):
    v &= (1 << 64) - 1
    out = bytearray()
    while True:
        b = v & 0x7F
        v >>= 7
        if v:
            out.append(b | 0x80)
        else:
            out.append(b)
            return bytes(out)


def _enc_ld(num, payload):
    return _enc_varint((num << 3) | 2) + _enc_varint(len(payload)) + payload


def encode_example(features):
    """dict name -> bytes | list of bytes | int sequence (np.integer dtype) | float sequence -> serialized Example
    (map entries in sorted key order, packed numeric lists: what protobuf's deterministic serialisation emits)."""
    entries = b""
    for key in sorted(features):
        v = features[key]
        if isinstance(v, (bytes, str)):
            v = [v]
        if len(v) and isinstance(v[0], (bytes, str)):
            lst = b"".join(_enc_ld(1, x if isinstance(x, bytes) else x.encode("utf-8")) for x in v)
            feat = _enc_ld(1, lst)
        else:
            a = np.asarray(v)
            if a.dtype.kind == "f":
                payload = a.astype("<f4").tobytes()
                feat = _enc_ld(2, _enc_ld(1, payload) if len(a) else b"")
            else:
                payload = b"".join(_enc_varint(int(x)) for x in a.reshape(-1))
                feat = _enc_ld(3, _enc_ld(1, payload) if len(a) else b"")
        entries += _enc_ld(1, _enc_ld(1, key.encode("utf-8")) + _enc_ld(2, feat))
    return _enc_ld(1, entries)


# ---------------------------------------------------------------------------------------------------------------------
# the reference's parse_fn + padded_batch
# ---------------------------------------------------------------------------------------------------------------------
def parse_sample(data, num_answers):
    """parse_fn of vqa/datasets/input_ops_vqa_tf_record_memft.py:28-62 on one serialized Example."""
    ex = parse_example(data)

    def scalar(name, kind, default):
        k, v = ex.get(name, ("none", None))
        if k != kind or v is None or len(v) == 0:
            if default is None:
                raise ValueError(f"feature {name!r} is required")   # FixedLenFeature without a default
            return default
        return v[0]

    def seq(name, kind, dtype):
        k, v = ex.get(name, ("none", None))
        return np.zeros(0, dtype) if k != kind or v is None else np.asarray(v, dtype)   # allow_missing=True

    ids = seq("answers/ids", "int64", np.int32)
    scores = seq("answers/scores", "float", np.float32)
    if len(ids) != len(scores):
        raise ValueError("answers/ids and answers/scores differ in length")
    target = np.zeros(num_answers, np.float32)        # tf.sparse_to_dense(ids, [num_answers], scores, 0)
    if len(ids):
        if ids.min() < 0 or ids.max() >= num_answers:
            raise ValueError("answer id out of range")
        target[ids] = scores
    return {
        "id": np.int64(scalar("qid", "int64", -1)),
        "image_id": scalar("image_id", "bytes", b""),
        "image_idx": np.int64(scalar("image_idx", "int64", -1)),
        "q_intseq": seq("q_intseq/list", "int64", np.int32),
        "q_intseq_len": np.int32(scalar("q_intseq/len", "int64", None)),
        "answer_target": target,
    }


def padded_batch(samples):
    """dataset.padded_batch (:64-73): q_intseq padded with 0 to the longest of the batch."""
    B = len(samples)
    T = max((len(s["q_intseq"]) for s in samples), default=0)
    q = np.zeros((B, T), np.int32)
    for i, s in enumerate(samples):
        q[i, :len(s["q_intseq"])] = s["q_intseq"]
    return {
        "id": np.array([s["id"] for s in samples], np.int64),
        "image_id": np.array([s["image_id"] for s in samples], dtype=object),
        "image_idx": np.array([s["image_idx"] for s in samples], np.int64),
        "q_intseq": q,
        "q_intseq_len": np.array([s["q_intseq_len"] for s in samples], np.int32),
        "answer_target": np.stack([s["answer_target"] for s in samples]) if B else np.zeros((0, 0), np.float32),
    }


def read_num_answers(tf_record_dir):
    """data_info.hdf5 ['data_info']['num_answers'] (input_ops...:13-15); .npz / .json mirrors where h5py is absent."""
    p = os.path.join(tf_record_dir, "data_info.hdf5")
    if os.path.exists(p):
        try:
            import h5py
            with h5py.File(p, "r") as f:
                return int(np.asarray(f["data_info"]["num_answers"]))
        except ImportError:
            from . import hdf5_min
            with hdf5_min.File(p) as f:
                return int(f["data_info/num_answers"])
    p = os.path.join(tf_record_dir, "data_info.npz")
    if os.path.exists(p):
        return int(np.load(p)["num_answers"])
    p = os.path.join(tf_record_dir, "data_info.json")
    if os.path.exists(p):
        import json
        with open(p) as f:
            return int(json.load(f)["num_answers"])
    raise ValueError(f"{tf_record_dir}: no readable data_info.{{hdf5,npz,json}} (h5py is needed for the .hdf5)")


def _interleave(files, cycle_length=10, verify_crc=True):
    """tf.contrib.data.parallel_interleave(TFRecordDataset, cycle_length=10, block_length=1), deterministic order:
    one record from each of the (up to) cycle_length open files in turn; an exhausted file is replaced by the next."""
    pending = list(files)
    open_its = []
    while pending and len(open_its) < cycle_length:
        open_its.append(read_records(pending.pop(0), verify_crc))
    i = 0
    while open_its:
        i %= len(open_its)
        try:
            yield next(open_its[i])
            i += 1
        except StopIteration:
            if pending:
                open_its[i] = read_records(pending.pop(0), verify_crc)
            else:
                open_its.pop(i)


def create(batch_size, tf_record_dir, split, is_train=True, scope="vqa_tf_record", shuffle=True, num_answers=None,
           seed=0, epochs=None, verify_crc=True):
    """Iterator of batch dicts with the keys / dtypes the reference's `create` returns as tensors. is_train: shuffle
    buffer of 3000 samples (:25-26), parsed samples cached in memory after the first pass (:75-76), repeated 1000
    times (:80-81; `epochs` overrides). The last batch of a pass may be smaller (padded_batch keeps the remainder)."""
    del scope
    if num_answers is None:
        num_answers = read_num_answers(tf_record_dir)
    files = sorted(glob.glob(os.path.join(tf_record_dir, split, f"{split}-*")))
    if not files:
        raise ValueError(f"no TFRecord shards match {os.path.join(tf_record_dir, split, split + '-*')}")
    rng = np.random.default_rng(seed)
    n_epochs = epochs if epochs is not None else (1000 if is_train else 1)

    def samples_once(cache):
        buf, size = [], 3000 if (is_train and shuffle) else 1
        source = cache if cache["done"] else None
        stream = (iter(source["items"]) if source else
                  (parse_sample(r, num_answers) for r in _interleave(files, 10, verify_crc)))
        for s in stream:
            if not cache["done"] and is_train:
                cache["items"].append(s)
            buf.append(s)
            if len(buf) >= size:
                yield buf.pop(int(rng.random() * len(buf)) if size > 1 else 0)
        while buf:
            yield buf.pop(int(rng.random() * len(buf)) if size > 1 else 0)
        cache["done"] = is_train

    def batches():
        # the reference shuffles BEFORE parse / batch / cache, so the cached order repeats every epoch (dataset.cache()
        # sits after padded_batch); the same here: batches of the first pass are cached and replayed
        cached = []
        cache = {"done": False, "items": []}
        for epoch in range(n_epochs):
            if epoch > 0 and is_train:
                for b in cached:
                    yield b
                continue
            cur = []
            for s in samples_once(cache):
                cur.append(s)
                if len(cur) == batch_size:
                    b = padded_batch(cur)
                    if is_train:
                        cached.append(b)
                    yield b
                    cur = []
            if cur:
                b = padded_batch(cur)
                if is_train:
                    cached.append(b)
                yield b

    return batches()


def write_shards(tf_record_dir, split, samples, num_answers, num_shards=2):
    """Write samples (dicts with qid, image_id, image_idx, q_intseq, answer ids / scores) as `split/split-XXXXX-of-
    YYYYY` TFRecord shards + data_info.npz, in the layout of generator_tf_record_memft_genome.py:184-195."""
    os.makedirs(os.path.join(tf_record_dir, split), exist_ok=True)
    per = (len(samples) + num_shards - 1) // num_shards
    for s in range(num_shards):
        chunk = samples[s * per:(s + 1) * per]
        payloads = [encode_example({
            "qid": np.asarray([x["qid"]], np.int64), "image_id": [x["image_id"]],
            "image_idx": np.asarray([x["image_idx"]], np.int64),
            "q_intseq/list": np.asarray(x["q_intseq"], np.int64),
            "q_intseq/len": np.asarray([len(x["q_intseq"])], np.int64),
            "answers/ids": np.asarray(x["answer_ids"], np.int64),
            "answers/scores": np.asarray(x["answer_scores"], np.float32),
        }) for x in chunk]
        write_records(os.path.join(tf_record_dir, split, f"{split}-{s:05d}-of-{num_shards:05d}"), payloads)
    np.savez(os.path.join(tf_record_dir, "data_info.npz"), num_answers=np.int64(num_answers))
