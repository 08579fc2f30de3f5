"""File contracts either side of the path: the exported vlmap word weights and the feature bank.

Reference behaviour mirrored (paths under /root/reference):
  vlmap/modules.py:589-614   WordWeightAnswer: column remap by answer STRING, absent -> weight 0, bias -100
  vlmap/modules.py:575-586   AnswerExistMask
  vlmap_memft/export_word_weights.py:60-83  weights.hdf5 {class_weights [J, A'], class_biases [A'], ...},
                                            vocab.pkl, answer_dict.pkl
  vqa/model_vlmap_answer.py:57-70           feature bank HDF5 {image_features, spatial_features,
                                            normal_boxes, num_boxes, data_info/{max_box_num, vfeat_dim}}
HDF5 is read through h5py when it is importable, else through hdf5_min (a pure-Python reader of the subset of the
format these files use); a `.npz` mirror with the same dataset names (weights.npz / *.npz next to the .hdf5 path) is
accepted too and takes precedence.
"""
import os
import pickle

import numpy as np

DEFAULT_BIAS = -100.0


def _load_pickle(path):
    with open(path, "rb") as f:
        return pickle.load(f, encoding="latin1")  # Python-2 cPickle files


def _read_datasets(path, names):
    """{name: array} from an HDF5 file or its .npz mirror."""
    npz = os.path.splitext(path)[0] + ".npz"
    if os.path.exists(npz):
        with np.load(npz) as z:
            return {n: np.asarray(z[n]) for n in names if n in z}
    try:
        import h5py
        opener = lambda p: h5py.File(p, "r")  # noqa: E731
    except ImportError:
        from . import hdf5_min                 # pure-Python reader of the subset these files use
        opener = hdf5_min.File
    out = {}
    with opener(path) as f:
        for n in names:
            for key in (n, "data_info/" + n):  # scalars such as max_box_num / vfeat_dim live in the data_info group
                if key in f:
                    out[n] = np.array(f[key])
                    break
    return out


def word_weight_answer(input_dim, answer_dict, word_weight_dir, weight_name="class_weights",
                       bias_name="class_biases", default_bias=DEFAULT_BIAS):
    """Initial value of WordWeightAnswer/fc/{weights [input_dim, A], biases [A]}."""
    A = len(answer_dict["vocab"])
    weights = np.zeros([input_dim, A], dtype=np.float32)
    biases = np.zeros([A], dtype=np.float32) + default_bias
    if word_weight_dir is not None:
        wdict = _load_pickle(os.path.join(word_weight_dir, "answer_dict.pkl"))
        d = _read_datasets(os.path.join(word_weight_dir, "weights.hdf5"), [weight_name, bias_name])
        aw, ab = d[weight_name], d[bias_name]
        for i, a in enumerate(answer_dict["vocab"]):
            j = wdict["dict"].get(a)
            if j is not None:
                weights[:, i] = aw[:, j]
                biases[i] = ab[j]
    return weights, biases


def answer_exist_mask(answer_dict, word_weight_dir=None):
    A = len(answer_dict["vocab"])
    if word_weight_dir is None:
        return np.ones([A], dtype=np.float32)
    wdict = _load_pickle(os.path.join(word_weight_dir, "answer_dict.pkl"))
    return np.array([1.0 if a in wdict["dict"] else 0.0 for a in answer_dict["vocab"]], dtype=np.float32)


def export_word_weights(out_dir, class_weights, class_biases, answer_vocab, extra=None, npz_mirror=False):
    """The exporter's contract (vlmap_memft/export_word_weights.py:60-83): `weights.hdf5` with the datasets
    class_weights [J, A'], class_biases [A'] (+ v_word, l_word, l_answer_word, or the v_/l_class_* sets of
    export_noc_word_weights.py:74-82, passed in `extra`) and the two pickles vocab.pkl / answer_dict.pkl. The HDF5 file
    is written by hdf5_min (contiguous datasets, the layout h5py's create_dataset(data=...) produces); npz_mirror also
    drops the .npz copy that _read_datasets prefers when present."""
    from . import hdf5_min
    os.makedirs(out_dir, exist_ok=True)
    arrays = {"class_weights": np.asarray(class_weights, np.float32),
              "class_biases": np.asarray(class_biases, np.float32)}
    arrays.update({k: np.asarray(v) for k, v in (extra or {}).items()})
    hdf5_min.write(os.path.join(out_dir, "weights.hdf5"), arrays)
    if npz_mirror:
        np.savez(os.path.join(out_dir, "weights.npz"), **arrays)
    adict = {"vocab": list(answer_vocab), "dict": {a: i for i, a in enumerate(answer_vocab)}}
    for name in ("answer_dict.pkl", "vocab.pkl"):
        with open(os.path.join(out_dir, name), "wb") as f:
            pickle.dump(adict, f, protocol=2)


def load_feature_bank(vfeat_path):
    d = _read_datasets(vfeat_path, ["image_features", "spatial_features", "normal_boxes", "num_boxes",
                                    "max_box_num", "vfeat_dim"])
    feats = d["image_features"]
    return {"features": feats, "spatials": d.get("spatial_features"), "normal_boxes": d.get("normal_boxes"),
            "num_boxes": d["num_boxes"], "max_box_num": int(d.get("max_box_num", feats.shape[1])),
            "vfeat_dim": int(d.get("vfeat_dim", feats.shape[2]))}
