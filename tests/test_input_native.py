"""Input side in native code (csrc/input_host.cu through vqa_transfer_externaldata_b200/input_native.py):
the C record indexer + Example parser must deliver the batches of the Python mirror (input_ops.create, itself pinned
against tensorboard's TF-team reader / writer in test_third_party_pin.py), in the same order, with the error behaviour
of the reference's parse_fn (vqa/datasets/input_ops_vqa_tf_record_memft.py:28-62). No GPU involved."""
import numpy as np
import pytest

from vqa_transfer_externaldata_b200 import input_native as IN
from vqa_transfer_externaldata_b200 import input_ops as IO
from vqa_transfer_externaldata_b200 import lib as L


def _samples(n, A, seed=3, tmax=14):
    rng = np.random.default_rng(seed)
    out = []
    for i in range(n):
        t = int(rng.integers(1, tmax + 1))
        k = int(rng.integers(0, 4))
        out.append({"qid": int(rng.integers(1, 1 << 40)), "image_id": b"COCO_%012d" % int(rng.integers(1 << 20)),
                    "image_idx": int(rng.integers(0, 4096)), "q_intseq": rng.integers(1, 8192, size=t).tolist(),
                    "answer_ids": rng.choice(A, size=k, replace=False).tolist(),
                    "answer_scores": rng.choice([0.3, 0.6, 0.9, 1.0], size=k).astype(np.float32).tolist()})
    return out


def _same(a, b):
    assert sorted(k for k in a if k not in ("answer_sparse", "num_answers")) == sorted(b)
    for k in b:
        if k == "image_id":
            assert list(a[k]) == list(b[k])
        else:
            assert a[k].dtype == b[k].dtype and a[k].shape == b[k].shape, k
            assert np.array_equal(a[k], b[k]), k


@pytest.mark.parametrize("is_train", [True, False])
def test_native_batches_equal_python_mirror(tmp_path, is_train):
    A = 300
    IO.write_shards(str(tmp_path), "train", _samples(3307, A), A, num_shards=13)   # > 10 shards: the interleave refills
    ref = IO.create(64, str(tmp_path), "train", is_train=is_train, seed=5, epochs=2 if is_train else None)
    got = IN.create(64, str(tmp_path), "train", is_train=is_train, seed=5, epochs=2 if is_train else None, dense_target=True)
    n = 0
    for a, b in zip(got, ref, strict=True):
        _same(a, b)
        rows, ids, scores = a["answer_sparse"]
        assert np.array_equal(IN.densify(a), b["answer_target"]) and len(rows) == int((b["answer_target"] != 0).sum())
        n += 1
    assert n == (2 if is_train else 1) * ((3307 + 63) // 64)


def test_parse_defaults_and_repeated_ids(tmp_path):
    """parse_fn defaults (qid -1, image_idx -1, image_id ''), allow_missing sequences, target[ids] = scores keeps the
    LAST score of a repeated id, unpacked numeric lists."""
    ex = [IO.encode_example({"q_intseq/len": np.array([0], np.int64)}),
          IO.encode_example({"q_intseq/len": np.array([2], np.int64), "q_intseq/list": np.array([7, 9], np.int64),
                             "answers/ids": np.array([4, 2, 4], np.int64),
                             "answers/scores": np.array([0.3, 0.6, 1.0], np.float32), "unknown": [b"zz"]})]
    p = tmp_path / "train"
    p.mkdir()
    IO.write_records(str(p / "train-00000-of-00001"), ex)
    b = next(IN.create(8, str(tmp_path), "train", is_train=False, num_answers=6, dense_target=True))
    assert b["id"].tolist() == [-1, -1] and b["image_idx"].tolist() == [-1, -1] and list(b["image_id"]) == [b"", b""]
    assert b["q_intseq"].tolist() == [[0, 0], [7, 9]] and b["q_intseq_len"].tolist() == [0, 2]
    assert b["answer_target"].tolist() == [[0] * 6, [0, 0, np.float32(0.6), 0, 1.0, 0]]
    rows, ids, scores = b["answer_sparse"]
    assert sorted(zip(rows.tolist(), ids.tolist())) == [(1, 2), (1, 4)]        # the earlier (4, 0.3) triple is dropped
    ref = next(IO.create(8, str(tmp_path), "train", is_train=False, num_answers=6))
    _same(b, ref)


def test_error_behaviour(tmp_path):
    d = tmp_path / "train"
    d.mkdir()
    f = str(d / "train-00000-of-00001")
    good = IO.encode_example({"q_intseq/len": np.array([1], np.int64), "q_intseq/list": np.array([3], np.int64)})
    # a required feature is missing (FixedLenFeature without a default)
    IO.write_records(f, [good, IO.encode_example({"qid": np.array([1], np.int64)})])
    with pytest.raises(L.VqaError, match="q_intseq/len"):
        next(IN.create(8, str(tmp_path), "train", is_train=False, num_answers=5))
    # answer id outside [0, num_answers)
    IO.write_records(f, [IO.encode_example({"q_intseq/len": np.array([0], np.int64), "answers/ids": np.array([5], np.int64),
                                            "answers/scores": np.array([1.0], np.float32)})])
    with pytest.raises(L.VqaError, match="out of range"):
        next(IN.create(8, str(tmp_path), "train", is_train=False, num_answers=5))
    # ids / scores of different length
    IO.write_records(f, [IO.encode_example({"q_intseq/len": np.array([0], np.int64), "answers/ids": np.array([1, 2], np.int64),
                                            "answers/scores": np.array([1.0], np.float32)})])
    with pytest.raises(L.VqaError, match="differ in length"):
        next(IN.create(8, str(tmp_path), "train", is_train=False, num_answers=5))
    # a flipped payload bit is caught by the record checksum; verify_crc=False reads past it into the parser
    IO.write_records(f, [good, good])
    raw = bytearray(open(f, "rb").read())
    raw[14] ^= 0x40
    open(f, "wb").write(bytes(raw))
    with pytest.raises(L.VqaError, match="corrupted record payload"):
        IN.create(8, str(tmp_path), "train", is_train=False, num_answers=5)
    # truncated file
    raw[14] ^= 0x40
    open(f, "wb").write(bytes(raw[:-3]))
    with pytest.raises(L.VqaError, match="truncated"):
        IN.create(8, str(tmp_path), "train", is_train=False, num_answers=5)
    # question longer than the row capacity
    IO.write_records(f, [IO.encode_example({"q_intseq/len": np.array([9], np.int64), "q_intseq/list": np.arange(1, 10, dtype=np.int64)})])
    with pytest.raises(L.VqaError, match="longer than"):
        next(IN.create(8, str(tmp_path), "train", is_train=False, num_answers=5, t_cap=8))


def test_record_index_matches_python_reader(tmp_path):
    payloads = [b"", b"a", bytes(range(256)) * 9, b"x" * 70000]
    p = str(tmp_path / "r")
    IO.write_records(p, payloads)
    s = IN.ShardSet([p])
    got = [s.data[0][int(o):int(o) + int(n)].tobytes() for o, n in zip(s.offsets[0], s.lengths[0])]
    assert got == payloads == list(IO.read_records(p))
