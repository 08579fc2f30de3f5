"""hdf5_min: the pure-Python HDF5 subset. Structures are pinned byte by byte against the HDF5 File Format
Specification (superblock v0, object header v1, symbol-table groups, layout v3); writer and reader against each other."""
import os
import struct

import numpy as np
import pytest

from vqa_transfer_externaldata_b200 import hdf5_min as H
from vqa_transfer_externaldata_b200 import wordweights as WW


def test_file_structure_follows_the_specification(tmp_path):
    p = str(tmp_path / "a.hdf5")
    H.write(p, {"x": np.arange(6, dtype=np.float32).reshape(2, 3)})
    b = open(p, "rb").read()
    # superblock version 0: signature, versions, 8-byte offsets and lengths, leaf K 4, internal K 16, base address 0,
    # end-of-file address = file size, root symbol table entry with cached B-tree / heap addresses (cache type 1)
    assert b[:8] == b"\x89HDF\r\n\x1a\n" and b[8] == 0 and b[13] == 8 and b[14] == 8
    assert struct.unpack_from("<HH", b, 16) == (4, 16)
    base, free, eof, drv = struct.unpack_from("<QQQQ", b, 24)
    assert base == 0 and free == H.UNDEF and eof == len(b) and drv == H.UNDEF
    name_off, root_hdr, cache, _ = struct.unpack_from("<QQII", b, 56)
    btree, heap = struct.unpack_from("<QQ", b, 80)
    assert name_off == 0 and cache == 1
    assert b[btree:btree + 4] == b"TREE" and b[btree + 4] == 0 and b[heap:heap + 4] == b"HEAP"
    # root object header, version 1: one symbol-table message (0x0011) carrying the same two addresses
    ver, _, nmsgs, refs, size = struct.unpack_from("<BBHII", b, root_hdr)
    assert (ver, nmsgs, refs) == (1, 1, 1)
    mtype, msize = struct.unpack_from("<HH", b, root_hdr + 16)
    assert mtype == 0x11 and msize == 16 and struct.unpack_from("<QQ", b, root_hdr + 24) == (btree, heap)
    # the B-tree's only child is a symbol table node whose single entry names "x" through the local heap
    snod = struct.unpack_from("<Q", b, btree + 32)[0]
    assert b[snod:snod + 4] == b"SNOD" and struct.unpack_from("<H", b, snod + 6)[0] == 1
    heap_data = struct.unpack_from("<Q", b, heap + 24)[0]
    link, hdr = struct.unpack_from("<QQ", b, snod + 8)
    assert b[heap_data + link:heap_data + link + 2] == b"x\x00"
    # dataset header: dataspace v1 (rank 2, dims 2 x 3), IEEE little-endian float32 datatype, contiguous layout v3
    msgs = {}
    pos = hdr + 16
    for _ in range(struct.unpack_from("<H", b, hdr + 2)[0]):
        t, s = struct.unpack_from("<HH", b, pos)
        msgs[t] = b[pos + 8:pos + 8 + s]
        pos += 8 + s
    assert msgs[1][:2] == bytes([1, 2]) and struct.unpack_from("<QQ", msgs[1], 8) == (2, 3)
    assert msgs[3][0] == 0x11 and struct.unpack_from("<I", msgs[3], 4)[0] == 4
    assert struct.unpack_from("<HHBBBBI", msgs[3], 8) == (0, 32, 23, 8, 0, 23, 127)
    assert msgs[8][:2] == bytes([3, 1])
    daddr, dsize = struct.unpack_from("<QQ", msgs[8], 2)
    assert dsize == 24 and np.array_equal(np.frombuffer(b, "<f4", 6, daddr), np.arange(6, dtype=np.float32))


def test_round_trip_of_the_reference_file_shapes(tmp_path):
    rng = np.random.default_rng(0)
    tree = {
        "image_features": rng.standard_normal((5, 4, 16)).astype(np.float32),
        "spatial_features": rng.standard_normal((5, 4, 6)).astype(np.float32),
        "normal_boxes": rng.uniform(size=(5, 4, 4)).astype(np.float32),
        "num_boxes": np.array([4, 1, 3, 2, 4], np.int32),
        "ids": np.arange(5, dtype=np.int64),
        "names": np.array([b"a", b"bcd", b"", b"zz", b"q"], dtype="S3"),
        "data_info": {"vfeat_dim": np.asarray(16, np.int32), "max_box_num": np.asarray(4, np.int32),
                      "pos": np.asarray(3.5, np.float64)},
    }
    tree.update({f"extra_{i:02d}": np.full(3, i, np.uint8) for i in range(20)})   # more than one symbol table node
    p = str(tmp_path / "bank.hdf5")
    H.write(p, tree)
    with H.File(p) as f:
        assert set(f.keys()) == set(tree) and f.keys("data_info") == ["max_box_num", "pos", "vfeat_dim"]
        for k, v in tree.items():
            if isinstance(v, dict):
                for k2, v2 in v.items():
                    got = f[f"{k}/{k2}"]
                    assert got.shape == () and got.dtype == v2.dtype and got == v2
            else:
                got = f[k]
                assert got.dtype == v.dtype and got.shape == v.shape and np.array_equal(got, v)
        assert "nope" not in f and "data_info/nope" not in f and f.get("nope") is None
        with pytest.raises(KeyError):
            f["data_info"]          # a group is not a dataset
    bank = WW.load_feature_bank(p)  # no h5py here: goes through hdf5_min
    assert bank["max_box_num"] == 4 and bank["vfeat_dim"] == 16
    assert np.array_equal(bank["features"], tree["image_features"]) and np.array_equal(bank["num_boxes"], tree["num_boxes"])


def test_chunked_deflate_shuffle(tmp_path):
    rng = np.random.default_rng(1)
    a = rng.integers(-1000, 1000, size=(7, 10)).astype(np.int32)
    w = rng.standard_normal((5, 9)).astype(np.float32)
    p = str(tmp_path / "c.hdf5")
    H.write(p, {"a": a, "w": w}, chunks={"a": (3, 4), "w": (5, 4)}, compress=True, shuffle=True)
    with H.File(p) as f:
        assert np.array_equal(f["a"], a) and np.array_equal(f["w"], w)   # ragged edge chunks, two filters undone in order
    H.write(p, {"a": a}, chunks={"a": (4, 4)})                            # chunked, no filter
    with H.File(p) as f:
        assert np.array_equal(f["a"], a)


def test_word_weights_and_data_info_through_hdf5(tmp_path):
    from vqa_transfer_externaldata_b200 import input_ops as IO
    import pickle
    d = str(tmp_path / "ww")
    os.makedirs(d)
    vocab = ["cat", "dog", "red"]
    cw = np.arange(12, dtype=np.float32).reshape(4, 3)
    H.write(os.path.join(d, "weights.hdf5"), {"class_weights": cw, "class_biases": np.array([1, 2, 3], np.float32)})
    for name in ("answer_dict.pkl", "vocab.pkl"):
        with open(os.path.join(d, name), "wb") as f:
            pickle.dump({"vocab": vocab, "dict": {a: i for i, a in enumerate(vocab)}}, f, protocol=2)
    mine = {"vocab": ["dog", "blue", "cat"]}
    w, b = WW.word_weight_answer(4, mine, d)
    assert np.array_equal(w[:, 0], cw[:, 1]) and np.array_equal(w[:, 2], cw[:, 0]) and not w[:, 1].any()
    assert b.tolist() == [2.0, -100.0, 1.0]                       # absent answer: weight 0, bias -100 (modules.py:600-614)
    H.write(str(tmp_path / "data_info.hdf5"), {"data_info": {"num_answers": np.asarray(3000, np.int32)}})
    assert IO.read_num_answers(str(tmp_path)) == 3000


def test_unsupported_files_say_so(tmp_path):
    p = str(tmp_path / "bad.hdf5")
    open(p, "wb").write(b"not hdf5 at all" * 10)
    with pytest.raises(ValueError):
        H.File(p)
    v2 = bytearray(H.SIGNATURE + bytes([2]) + bytes(100))
    open(p, "wb").write(bytes(v2))
    with pytest.raises(NotImplementedError):
        H.File(p)
