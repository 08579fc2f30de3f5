"""Input side: TFRecord framing, tf.train.Example wire format, and the reference's parse_fn / padded_batch
(vqa/datasets/input_ops_vqa_tf_record_memft.py:17-82), pinned by known answers derived from the published formats."""
import os
import struct

import numpy as np
import pytest

from vqa_transfer_externaldata_b200 import input_ops as IO


def test_crc32c_known_answers():
    # RFC 3720 B.4 / the standard CRC-32C check value
    assert IO.crc32c(b"123456789") == 0xE3069283
    assert IO.crc32c(b"") == 0
    assert IO.crc32c(bytes(32)) == 0x8A9136AA           # 32 zero bytes (RFC 3720 B.4)
    assert IO.crc32c(bytes([0xFF] * 32)) == 0x62A8AB43  # 32 0xFF bytes
    assert IO.crc32c(bytes(range(32))) == 0x46DD794E    # ascending
    # the pure-Python table walk and the C library agree
    c = 0xFFFFFFFF
    for b in b"The quick brown fox jumps over the lazy dog":
        c = IO._TABLE[(c ^ b) & 0xFF] ^ (c >> 8)
    assert IO.crc32c(b"The quick brown fox jumps over the lazy dog") == (c ^ 0xFFFFFFFF) == 0x22620404


def test_example_wire_format_known_answer():
    """Bytes written out by hand from the protobuf encoding rules: Example{features{feature{'a': int64_list [1, 300]}}}"""
    int64_list = bytes([0x0A, 0x03, 0x01, 0xAC, 0x02])             # field 1, packed: 1, 300
    feature = bytes([0x1A, len(int64_list)]) + int64_list           # Feature.int64_list = field 3
    entry = bytes([0x0A, 0x01]) + b"a" + bytes([0x12, len(feature)]) + feature
    features = bytes([0x0A, len(entry)]) + entry
    example = bytes([0x0A, len(features)]) + features
    assert IO.encode_example({"a": np.array([1, 300], np.int64)}) == example
    kind, val = IO.parse_example(example)["a"]
    assert kind == "int64" and val.tolist() == [1, 300]
    # negative values, float and bytes lists
    ex = IO.encode_example({"f": np.array([0.5, -2.0], np.float32), "s": [b"xy", b""], "n": np.array([-1], np.int64)})
    got = IO.parse_example(ex)
    assert got["f"][0] == "float" and got["f"][1].tolist() == [0.5, -2.0]
    assert got["s"] == ("bytes", [b"xy", b""])
    assert got["n"][1].tolist() == [-1]
    # a FloatList / Int64List written element by element (not packed) parses the same
    fl = bytes([0x0D]) + struct.pack("<f", 1.5) + bytes([0x0D]) + struct.pack("<f", 2.5)
    feat = bytes([0x12, len(fl)]) + fl
    ent = bytes([0x0A, 0x01]) + b"u" + bytes([0x12, len(feat)]) + feat
    fs = bytes([0x0A, len(ent)]) + ent
    assert IO.parse_example(bytes([0x0A, len(fs)]) + fs)["u"][1].tolist() == [1.5, 2.5]


def test_record_framing_and_corruption(tmp_path):
    p = str(tmp_path / "r")
    IO.write_records(p, [b"hello", b"", b"x" * 1000])
    assert list(IO.read_records(p)) == [b"hello", b"", b"x" * 1000]
    raw = open(p, "rb").read()
    # first record: length 5, masked crc of the length bytes, payload, masked crc of the payload
    assert struct.unpack("<Q", raw[:8])[0] == 5 and raw[12:17] == b"hello"
    c = IO.crc32c(b"hello")
    assert struct.unpack("<I", raw[17:21])[0] == (((c >> 15) | (c << 17)) + 0xA282EAD8) & 0xFFFFFFFF
    bad = bytearray(raw)
    bad[13] ^= 1
    open(p, "wb").write(bytes(bad))
    with pytest.raises(ValueError):
        list(IO.read_records(p))
    assert list(IO.read_records(p, verify_crc=False))[0] == b"hdllo"
    open(p, "wb").write(raw[:-3])
    with pytest.raises(ValueError):
        list(IO.read_records(p))


def _samples(n, A=40, seed=0):
    rng = np.random.default_rng(seed)
    out = []
    for i in range(n):
        k = int(rng.integers(0, 4))                     # some questions have no answer in the vocabulary
        out.append({"qid": 1000 + i, "image_id": f"img{i % 7}".encode(), "image_idx": i % 7,
                    "q_intseq": rng.integers(1, 90, size=int(rng.integers(1, 15))).astype(np.int32),
                    "answer_ids": rng.choice(A, size=k, replace=False).astype(np.int64),
                    "answer_scores": rng.choice([0.3, 0.6, 0.9, 1.0], size=k).astype(np.float32)})
    return out


def test_create_reads_shards_like_the_reference_pipeline(tmp_path):
    A, n = 40, 53
    samples = _samples(n, A)
    IO.write_shards(str(tmp_path), "val", samples, A, num_shards=3)
    assert IO.read_num_answers(str(tmp_path)) == A
    batches = list(IO.create(16, str(tmp_path), "val", is_train=False))
    assert [len(b["id"]) for b in batches] == [16, 16, 16, 5]          # padded_batch keeps the remainder
    by_qid = {s["qid"]: s for s in samples}
    seen = []
    for b in batches:
        assert b["id"].dtype == np.int64 and b["image_idx"].dtype == np.int64
        assert b["q_intseq"].dtype == np.int32 and b["q_intseq_len"].dtype == np.int32
        assert b["answer_target"].dtype == np.float32 and b["answer_target"].shape == (len(b["id"]), A)
        assert b["q_intseq"].shape[1] == b["q_intseq_len"].max()         # padded to the longest of THIS batch
        for i, qid in enumerate(b["id"]):
            s = by_qid[int(qid)]
            seen.append(int(qid))
            L = len(s["q_intseq"])
            assert b["q_intseq_len"][i] == L and b["q_intseq"][i, :L].tolist() == s["q_intseq"].tolist()
            assert not b["q_intseq"][i, L:].any()                        # pad id 0
            t = np.zeros(A, np.float32)
            t[s["answer_ids"]] = s["answer_scores"]
            assert np.array_equal(b["answer_target"][i], t)              # sparse_to_dense
            assert b["image_id"][i] == s["image_id"] and b["image_idx"][i] == s["image_idx"]
    assert sorted(seen) == sorted(by_qid)
    # parallel_interleave(cycle_length=10, block_length=1): one record from each shard in turn
    shard0, shard1, shard2 = samples[0]["qid"], samples[18]["qid"], samples[36]["qid"]
    assert seen[:3] == [shard0, shard1, shard2]


def test_train_pipeline_shuffles_caches_and_repeats(tmp_path):
    A, n = 40, 50
    IO.write_shards(str(tmp_path), "train", _samples(n, A, seed=1), A, num_shards=2)
    it = IO.create(8, str(tmp_path), "train", is_train=True, shuffle=True, seed=3, epochs=2)
    batches = list(it)
    assert len(batches) == 2 * 7
    first = np.concatenate([b["id"] for b in batches[:7]])
    second = np.concatenate([b["id"] for b in batches[7:]])
    assert sorted(first.tolist()) == list(range(1000, 1000 + n))
    assert first.tolist() != list(range(1000, 1000 + n))                # shuffled
    assert np.array_equal(first, second)                               # dataset.cache() replays the first pass
    # missing scalar without a default is an error (FixedLenFeature((), tf.int64) for q_intseq/len)
    with pytest.raises(ValueError):
        IO.parse_sample(IO.encode_example({"qid": np.array([1], np.int64)}), A)
    d = IO.parse_sample(IO.encode_example({"q_intseq/len": np.array([0], np.int64)}), A)
    assert d["id"] == -1 and d["image_id"] == b"" and d["image_idx"] == -1 and len(d["q_intseq"]) == 0
