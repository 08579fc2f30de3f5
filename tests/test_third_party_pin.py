"""f3 / f4 cross-checked against INDEPENDENT implementations that ship in this image (no TensorFlow needed):

  * TFRecord framing + masked CRC-32C: tensorboard's TF-team code -- `tensorboard.summary.writer.record_writer.RecordWriter`
    (writer) and `tensorboard.compat.tensorflow_stub.pywrap_tensorflow.PyRecordReader_New` / `masked_crc32c` (reader):
    our reader opens files THEIR writer wrote, THEIR reader opens files our writer wrote, checksums agree byte for byte.
  * tf.train.Example wire format: messages built at run time by the protobuf RUNTIME (google.protobuf) from the published
    schema of tensorflow/core/example/{example,feature}.proto (field numbers below): our hand-rolled encoder's bytes
    parse there and its serializer's bytes parse here, packed and unpacked repeated fields, negative int64s included.
  * tensor-bundle protos: the `BundleEntryProto.shape` / `.dtype` and `BundleHeaderProto.version` fields of our checkpoint
    index against tensorboard.compat.proto.{tensor_shape_pb2, types_pb2, versions_pb2} (generated from TensorFlow's own
    .proto files).
  * weights.hdf5 written by wordweights.export_word_weights is read back through the HDF5 subset reader.
"""
import os
import struct

import numpy as np
import pytest

from vqa_transfer_externaldata_b200 import input_ops as IO
from vqa_transfer_externaldata_b200 import tf_bundle as TB

tb_writer = pytest.importorskip("tensorboard.summary.writer.record_writer")
pw = pytest.importorskip("tensorboard.compat.tensorflow_stub.pywrap_tensorflow")


def _payloads():
    rng = np.random.default_rng(0)
    return [b"", b"x", bytes(rng.integers(0, 256, size=1000, dtype=np.uint8)),
            bytes(rng.integers(0, 256, size=70001, dtype=np.uint8))]


def test_masked_crc_matches_tensorboard():
    for data in _payloads() + [b"123456789", struct.pack("<Q", 12345)]:
        assert IO.masked_crc(data) == int(pw.masked_crc32c(data))
        assert IO.crc32c(data) == int(pw.crc32c(data))


def test_our_reader_opens_tensorboard_written_records(tmp_path):
    path = str(tmp_path / "tb.tfrecord")
    with open(path, "wb") as f:
        w = tb_writer.RecordWriter(f)
        for p in _payloads():
            w.write(p)
        w.flush()
    assert list(IO.read_records(path)) == _payloads()


def test_tensorboard_reader_opens_our_records(tmp_path):
    path = str(tmp_path / "ours.tfrecord")
    IO.write_records(path, _payloads())
    r = pw.PyRecordReader_New(path)
    got = []
    while True:
        try:
            r.GetNext()
        except Exception:  # noqa: BLE001 -- the stub raises OutOfRangeError at the end of the file
            break
        got.append(bytes(r.record()))
    assert got == _payloads()
    # and byte for byte the same file as their writer produces
    path2 = str(tmp_path / "tb.tfrecord")
    with open(path2, "wb") as f:
        w = tb_writer.RecordWriter(f)
        for p in _payloads():
            w.write(p)
        w.flush()
    assert open(path, "rb").read() == open(path2, "rb").read()


# ---- tf.train.Example through the protobuf runtime -------------------------------------------------------------------
def _example_classes():
    """Message classes for tensorflow/core/example/{feature,example}.proto built by the protobuf runtime.
    BytesList{1: repeated bytes}, FloatList{1: repeated float [packed]}, Int64List{1: repeated int64 [packed]},
    Feature{oneof kind {1: BytesList, 2: FloatList, 3: Int64List}}, Features{1: map<string, Feature>},
    Example{1: Features}."""
    from google.protobuf import descriptor_pb2, descriptor_pool, message_factory
    fd = descriptor_pb2.FileDescriptorProto()
    fd.name, fd.package, fd.syntax = "vqa_test_example.proto", "vqa_test", "proto3"
    T = descriptor_pb2.FieldDescriptorProto

    def msg(name):
        m = fd.message_type.add()
        m.name = name
        return m

    def field(m, name, number, ftype, label=T.LABEL_OPTIONAL, type_name=None, packed=None):
        f = m.field.add()
        f.name, f.number, f.type, f.label = name, number, ftype, label
        if type_name:
            f.type_name = type_name
        if packed is not None:
            f.options.packed = packed
        return f

    field(msg("BytesList"), "value", 1, T.TYPE_BYTES, T.LABEL_REPEATED)
    field(msg("FloatList"), "value", 1, T.TYPE_FLOAT, T.LABEL_REPEATED, packed=True)
    field(msg("Int64List"), "value", 1, T.TYPE_INT64, T.LABEL_REPEATED, packed=True)
    feat = msg("Feature")
    feat.oneof_decl.add().name = "kind"
    for name, num, tn in (("bytes_list", 1, "BytesList"), ("float_list", 2, "FloatList"), ("int64_list", 3, "Int64List")):
        f = field(feat, name, num, T.TYPE_MESSAGE, type_name=".vqa_test." + tn)
        f.oneof_index = 0
    feats = msg("Features")
    entry = feats.nested_type.add()
    entry.name = "FeatureEntry"
    entry.options.map_entry = True
    field(entry, "key", 1, T.TYPE_STRING)
    field(entry, "value", 2, T.TYPE_MESSAGE, type_name=".vqa_test.Feature")
    field(feats, "feature", 1, T.TYPE_MESSAGE, T.LABEL_REPEATED, type_name=".vqa_test.Features.FeatureEntry")
    field(msg("Example"), "features", 1, T.TYPE_MESSAGE, type_name=".vqa_test.Features")
    pool = descriptor_pool.DescriptorPool()
    pool.Add(fd)
    return message_factory.GetMessageClass(pool.FindMessageTypeByName("vqa_test.Example"))


SAMPLE = {"qid": np.array([262148000], np.int64), "image_id": b"COCO_train2014_000000262148",
          "image_idx": np.array([-1], np.int64), "q_intseq/list": np.array([5, 17, 300, 70000, 2], np.int64),
          "q_intseq/len": np.array([5], np.int64), "answers/ids": np.array([3, 2999], np.int64),
          "answers/scores": np.array([0.3, 1.0], np.float32)}


def test_our_example_bytes_parse_in_the_protobuf_runtime():
    Example = _example_classes()
    ex = Example()
    ex.ParseFromString(IO.encode_example(SAMPLE))
    f = ex.features.feature
    assert list(f["q_intseq/list"].int64_list.value) == [5, 17, 300, 70000, 2]
    assert list(f["image_idx"].int64_list.value) == [-1]
    assert list(f["image_id"].bytes_list.value) == [b"COCO_train2014_000000262148"]
    assert np.allclose(list(f["answers/scores"].float_list.value), [0.3, 1.0])
    assert sorted(f.keys()) == sorted(SAMPLE)
    # the runtime's own serialisation of the same message: the same bytes up to the (semantically free) order of the map
    # entries -- compare the multiset of serialized entries
    def entries(buf):
        (feats,) = [v for num, wt, v in IO._fields(memoryview(buf)) if num == 1]
        return sorted(bytes(v) for num, wt, v in IO._fields(feats) if num == 1)
    theirs = ex.SerializeToString(deterministic=True)
    assert len(theirs) == len(IO.encode_example(SAMPLE))
    assert entries(theirs) == entries(IO.encode_example(SAMPLE))


def test_protobuf_runtime_bytes_parse_here():
    Example = _example_classes()
    ex = Example()
    f = ex.features.feature
    f["qid"].int64_list.value.append(7)
    f["image_id"].bytes_list.value.append(b"img")
    f["image_idx"].int64_list.value.append(-1)
    f["q_intseq/list"].int64_list.value.extend([9, 8, 1 << 40])
    f["q_intseq/len"].int64_list.value.append(3)
    f["answers/ids"].int64_list.value.extend([1, 4])
    f["answers/scores"].float_list.value.extend([0.6, 0.9])
    got = IO.parse_example(ex.SerializeToString())
    assert got["image_idx"][1].tolist() == [-1]
    assert got["q_intseq/list"][1].tolist() == [9, 8, 1 << 40]
    assert got["image_id"][1] == [b"img"]
    assert np.allclose(got["answers/scores"][1], [0.6, 0.9])
    s = IO.parse_sample(ex.SerializeToString(), 6)
    assert s["answer_target"].tolist() == pytest.approx([0, 0.6, 0, 0, 0.9, 0])
    assert s["q_intseq_len"] == 3 and s["image_idx"] == -1


# ---- tensor-bundle protos ----------------------------------------------------------------------------------------------
def test_bundle_entry_fields_match_tensorflow_protos():
    from tensorboard.compat.proto import tensor_shape_pb2, types_pb2, versions_pb2
    assert TB._DT[np.dtype(np.float32)] == types_pb2.DT_FLOAT
    assert TB._DT[np.dtype(np.int64)] == types_pb2.DT_INT64
    assert TB._DT[np.dtype(np.int32)] == types_pb2.DT_INT32
    for shape in ((), (7,), (1324, 2048), (2, 3, 4)):
        ent = TB._encode_entry(types_pb2.DT_FLOAT, shape, 128, 4 * int(np.prod(shape)), 0xDEADBEEF)
        # field 2 of BundleEntryProto is a TensorShapeProto: take its bytes and let TensorFlow's generated class parse them
        fields = {num: v for num, wt, v in IO._fields(memoryview(ent))}
        proto = tensor_shape_pb2.TensorShapeProto()
        proto.ParseFromString(bytes(fields[2]))
        assert [d.size for d in proto.dim] == list(shape) and not proto.unknown_rank
        want = tensor_shape_pb2.TensorShapeProto(dim=[tensor_shape_pb2.TensorShapeProto.Dim(size=n) for n in shape])
        assert bytes(fields[2]) == want.SerializeToString()
        assert TB._decode_entry(ent)["shape"] == list(shape)


def test_bundle_header_version_parses_as_versiondef(tmp_path):
    from tensorboard.compat.proto import versions_pb2
    prefix = str(tmp_path / "model-1")
    TB.write_bundle(prefix, {"a": np.arange(6, dtype=np.float32).reshape(2, 3), "global_step": np.asarray(1, np.int64)})
    entries = dict(TB.read_table(prefix + ".index"))
    header = {num: v for num, wt, v in IO._fields(memoryview(entries[b""]))}
    assert header[1] == 1                                  # num_shards
    v = versions_pb2.VersionDef()
    v.ParseFromString(bytes(header[3]))                    # BundleHeaderProto.version
    assert v.producer == 1 and v.min_consumer == 0


# ---- weights.hdf5 ------------------------------------------------------------------------------------------------------
def test_export_word_weights_writes_real_hdf5(tmp_path):
    from vqa_transfer_externaldata_b200 import hdf5_min, wordweights as WW
    rng = np.random.default_rng(3)
    vocab = [f"ans{i}" for i in range(9)]
    cw = rng.standard_normal((16, 9)).astype(np.float32)
    cb = rng.standard_normal(9).astype(np.float32)
    extra = {"v_word": rng.standard_normal((5, 8)).astype(np.float32), "l_word": rng.standard_normal((5, 8)).astype(np.float32),
             "l_answer_word": rng.standard_normal((9, 8)).astype(np.float32)}
    out = str(tmp_path / "word_weights_model-4801")
    WW.export_word_weights(out, cw, cb, vocab, extra=extra)
    path = os.path.join(out, "weights.hdf5")
    assert os.path.exists(path) and open(path, "rb").read(8) == b"\x89HDF\r\n\x1a\n"
    with hdf5_min.File(path) as f:
        assert sorted(f.keys()) == sorted(["class_weights", "class_biases", "v_word", "l_word", "l_answer_word"])
        assert np.array_equal(np.array(f["class_weights"]), cw)
        assert np.array_equal(np.array(f["l_answer_word"]), extra["l_answer_word"])
    os.remove(os.path.join(out, "weights.npz")) if os.path.exists(os.path.join(out, "weights.npz")) else None
    adict = {"vocab": ["ans3", "zzz", "ans0"]}
    w, b = WW.word_weight_answer(16, adict, out)
    assert np.array_equal(w[:, 0], cw[:, 3]) and np.array_equal(w[:, 2], cw[:, 0]) and np.all(w[:, 1] == 0)
    assert b[1] == -100.0 and b[0] == cb[3]
