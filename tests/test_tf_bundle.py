"""TensorFlow checkpoint bundles without TensorFlow: the sorted-string-table container and the bundle
protos, pinned by known answers derived from the published format and by writer / reader round trips."""
import struct

import numpy as np
import pytest

from vqa_transfer_externaldata_b200 import tf_bundle as TB
from vqa_transfer_externaldata_b200.input_ops import crc32c


def test_block_format_known_answer():
    """A block written out by hand: 'apple' -> '1', 'apply' -> '22' (shares 'appl'), one restart at offset 0."""
    block = bytes([0, 5, 1]) + b"apple" + b"1" + bytes([4, 1, 2]) + b"y" + b"22" + struct.pack("<I", 0) + struct.pack("<I", 1)
    assert TB._parse_block(block) == [(b"apple", b"1"), (b"apply", b"22")]
    bb = TB._BlockBuilder()
    bb.add(b"apple", b"1")
    bb.add(b"apply", b"22")
    assert bb.finish() == block


def test_table_round_trip_footer_and_checksums(tmp_path):
    p = str(tmp_path / "t.index")
    rng = np.random.default_rng(0)
    items = [(f"scope_{i // 7}/layer_{i}/weights".encode(), rng.bytes(int(rng.integers(0, 300)))) for i in range(200)]
    TB.write_table(p, items)
    assert TB.read_table(p) == sorted(items)          # several 4 KB blocks, restarts every 16 entries
    raw = open(p, "rb").read()
    assert struct.unpack("<Q", raw[-8:])[0] == 0xDB4775248B80FB57 and len(raw) >= 48
    # first data block: trailer = type 0 + masked crc32c(block + type)
    moff, pos = TB._varint(raw[-48:], 0)
    first_index_entry = TB.read_table(p)[0]
    assert first_index_entry[0] == sorted(items)[0][0]
    bad = bytearray(raw)
    bad[10] ^= 0x40
    open(p, "wb").write(bytes(bad))
    with pytest.raises(ValueError):
        TB.read_table(p)
    bad = bytearray(raw)
    bad[-1] ^= 1
    open(p, "wb").write(bytes(bad))
    with pytest.raises(ValueError):
        TB.read_table(p)
    del moff, pos


def test_bundle_entry_proto_known_answer():
    """BundleEntryProto{dtype: DT_FLOAT, shape: [2, 3], offset: 128, size: 24, crc32c: 0x01020304} by the encoding rules."""
    want = (bytes([0x08, 0x01])                                                # 1: dtype = 1
            + bytes([0x12, 0x08, 0x12, 0x02, 0x08, 0x02, 0x12, 0x02, 0x08, 0x03])   # 2: shape {dim{size 2} dim{size 3}}
            + bytes([0x20, 0x80, 0x01])                                        # 4: offset = 128
            + bytes([0x28, 0x18])                                              # 5: size = 24
            + bytes([0x35, 0x04, 0x03, 0x02, 0x01]))                           # 6: fixed32
    assert TB._encode_entry(TB.DT_FLOAT, (2, 3), 128, 24, 0x01020304) == want
    ent = TB._decode_entry(want)
    assert ent == {"dtype": 1, "shape": [2, 3], "shard_id": 0, "offset": 128, "size": 24, "crc32c": 0x01020304}


def test_bundle_round_trip_and_corruption(tmp_path):
    rng = np.random.default_rng(1)
    tensors = {
        "v_linear_v/fc/weights": rng.standard_normal((64, 32)).astype(np.float32),
        "v_linear_v/fc/biases": rng.standard_normal(32).astype(np.float32),
        "encode_L/rnn/gru_cell/gates/kernel": rng.standard_normal((50, 64)).astype(np.float32),
        "global_step": np.asarray(12000, np.int64),
        "some/int32": np.arange(7, dtype=np.int32),
    }
    prefix = str(tmp_path / "model-12000")
    TB.write_bundle(prefix, tensors)
    got = TB.read_bundle(prefix)
    assert set(got) == set(tensors)
    for k, v in tensors.items():
        assert got[k].dtype == v.dtype and got[k].shape == v.shape and np.array_equal(got[k], v)
    assert list(TB.read_bundle(prefix, names={"global_step"})) == ["global_step"]
    # the header entry: key "", num_shards 1, little endian, version producer 1
    header = dict(TB.read_table(prefix + ".index"))[b""]
    assert header == bytes([0x08, 0x01, 0x1A, 0x02, 0x08, 0x01])
    # tensor bytes are raw little endian at the recorded offset, with the masked CRC-32C of exactly those bytes
    ent = TB._decode_entry(dict(TB.read_table(prefix + ".index"))[b"v_linear_v/fc/biases"])
    raw = open(prefix + ".data-00000-of-00001", "rb").read()[ent["offset"]:ent["offset"] + ent["size"]]
    assert raw == tensors["v_linear_v/fc/biases"].tobytes()
    c = crc32c(raw)
    assert ent["crc32c"] == (((c >> 15) | (c << 17)) + 0xA282EAD8) & 0xFFFFFFFF
    data = bytearray(open(prefix + ".data-00000-of-00001", "rb").read())
    data[5] ^= 1
    open(prefix + ".data-00000-of-00001", "wb").write(bytes(data))
    with pytest.raises(ValueError):
        TB.read_bundle(prefix)
    assert set(TB.read_bundle(prefix, verify=False)) == set(tensors)
