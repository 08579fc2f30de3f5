"""Input side of the path on the C library: the same batches as input_ops.create -- and, under the same
seed, the SAME ORDER -- but framing, checksum, tf.train.Example parsing, parse_fn defaults and padded_batch run in
native code (csrc/input_host.cu: vqa_tfrecord_index_host, vqa_parse_examples_host), one call per batch, and the dense
[B, num_answers] soft-score target is never built on the host: a batch carries the (row, answer id, score) triples
(`answer_sparse`) and Engine.stage_batch scatters them on the device (vqa_densify_targets).

Mirrors vqa/datasets/input_ops_vqa_tf_record_memft.py:6-82 (create: interleave 10 files, shuffle buffer 3000, parse_fn,
padded_batch, cache, repeat). Measured in profiles/r02_input_pipeline.md: the pure-Python mirror delivers ~10^4
samples/s, this module > 10^6 per host thread; a cfg1 step consumes 5 x 10^5.
"""
import ctypes as C
import glob
import os

import numpy as np

from . import input_ops as IO
from . import lib as L

_U8P = C.POINTER(C.c_uint8)


def _ptr(a):
    return C.c_void_p(a.ctypes.data)


class ShardSet:
    """The TFRecord shards of one split, held in host memory and indexed by the C library."""

    def __init__(self, files, verify_crc=True):
        self.lib = L.load()
        self.files = list(files)
        self.data, self.offsets, self.lengths = [], [], []
        for path in self.files:
            buf = np.fromfile(path, dtype=np.uint8)
            n = C.c_int64(0)
            cap = max(16, buf.size // 64)
            while True:
                off = np.zeros(cap, np.uint64)
                ln = np.zeros(cap, np.uint64)
                if buf.size == 0:
                    break
                L.check(self.lib.vqa_tfrecord_index_host(_ptr(buf), C.c_uint64(buf.size), C.c_int32(1 if verify_crc else 0),
                                                         _ptr(off), _ptr(ln), C.c_int64(cap), C.byref(n)))
                if n.value <= cap:
                    break
                cap = int(n.value)
            self.data.append(buf)
            self.offsets.append(off[:n.value].copy())
            self.lengths.append(ln[:n.value].copy())

    def interleaved(self, cycle_length=10):
        """Record order of tf.contrib.data.parallel_interleave(cycle_length, block_length=1) (input_ops._interleave):
        arrays (file, record) of every record."""
        counts = [len(o) for o in self.offsets]
        pending = list(range(len(self.files)))
        open_f, pos = [], []
        while pending and len(open_f) < cycle_length:
            open_f.append(pending.pop(0))
            pos.append(0)
        order_f, order_r = [], []
        i = 0
        while open_f:
            i %= len(open_f)
            f = open_f[i]
            if pos[i] < counts[f]:
                order_f.append(f)
                order_r.append(pos[i])
                pos[i] += 1
                i += 1
            elif pending:
                open_f[i] = pending.pop(0)
                pos[i] = 0
            else:
                open_f.pop(i)
                pos.pop(i)
        return np.asarray(order_f, np.int64), np.asarray(order_r, np.int64)

    def parse(self, file_idx, rec_idx, num_answers, t_cap=64, want_image_id=True):
        """parse_fn + padded_batch of the given records -> batch dict (sparse soft-score target)."""
        n = len(file_idx)
        ptrs = np.empty(n, np.uint64)
        lens = np.empty(n, np.uint64)
        for k in range(len(self.files)):
            sel = np.nonzero(file_idx == k)[0]
            if sel.size:
                ptrs[sel] = np.uint64(self.data[k].ctypes.data) + self.offsets[k][rec_idx[sel]]
                lens[sel] = self.lengths[k][rec_idx[sel]]
        qid = np.empty(n, np.int64)
        iidx = np.empty(n, np.int64)
        q = np.empty((n, t_cap), np.int32)
        qlen = np.empty(n, np.int32)
        cap = 16 * max(n, 1)
        rows = np.empty(cap, np.int32)
        ids = np.empty(cap, np.int32)
        scores = np.empty(cap, np.float32)
        ioff = np.zeros(n, np.uint32)
        ilen = np.zeros(n, np.uint32)
        t_long, na = C.c_int32(0), C.c_int32(0)
        L.check(self.lib.vqa_parse_examples_host(_ptr(ptrs), _ptr(lens), C.c_int32(n), C.c_int32(num_answers),
                                                 C.c_int32(t_cap), _ptr(qid), _ptr(iidx), _ptr(q), _ptr(qlen),
                                                 C.byref(t_long), _ptr(rows), _ptr(ids), _ptr(scores), C.c_int32(cap),
                                                 C.byref(na), _ptr(ioff) if want_image_id else C.c_void_p(0),
                                                 _ptr(ilen) if want_image_id else C.c_void_p(0)))
        T = int(t_long.value)
        batch = {"id": qid, "image_idx": iidx, "q_intseq": np.ascontiguousarray(q[:, :T]), "q_intseq_len": qlen,
                 "answer_sparse": (rows[:na.value].copy(), ids[:na.value].copy(), scores[:na.value].copy()),
                 "num_answers": int(num_answers)}
        if want_image_id:
            image_id = np.empty(n, dtype=object)
            for i in range(n):
                k = int(file_idx[i])
                base = int(self.offsets[k][rec_idx[i]]) + int(ioff[i])
                image_id[i] = self.data[k][base:base + int(ilen[i])].tobytes() if ilen[i] else b""
            batch["image_id"] = image_id
        return batch


def densify(batch):
    """Host-side tf.sparse_to_dense of a batch's triples (tests, CPU consumers); the GPU path scatters on the device."""
    rows, ids, scores = batch["answer_sparse"]
    t = np.zeros((len(batch["q_intseq_len"]), batch["num_answers"]), np.float32)
    t[rows, ids] = scores
    return t


def create(batch_size, tf_record_dir, split, is_train=True, scope="vqa_tf_record", shuffle=True, num_answers=None,
           seed=0, epochs=None, verify_crc=True, dense_target=False, t_cap=64, want_image_id=True):
    """Same contract and (for a given seed) the same batches in the same order as input_ops.create; `answer_target` is
    present only with dense_target=True, `answer_sparse` always."""
    del scope
    if num_answers is None:
        num_answers = IO.read_num_answers(tf_record_dir)
    files = sorted(glob.glob(os.path.join(tf_record_dir, split, f"{split}-*")))
    if not files:
        raise ValueError(f"no TFRecord shards match {os.path.join(tf_record_dir, split, split + '-*')}")
    shards = ShardSet(files, verify_crc)
    of, orr = shards.interleaved(10)
    rng = np.random.default_rng(seed)
    n_epochs = epochs if epochs is not None else (1000 if is_train else 1)

    def shuffled_order():
        # the shuffle buffer of input_ops.create (3000 samples), on record numbers
        size = 3000 if (is_train and shuffle) else 1
        if size == 1:
            return np.arange(len(of))
        # (one uniform double per pop, as input_ops.create draws them; drawn in bulk: the stream is the same)
        n = len(of)
        u = rng.random(n).tolist()
        buf, out = list(range(min(size - 1, n))), []
        for k in range(len(buf), n):
            buf.append(k)
            out.append(buf.pop(int(u[len(out)] * size)))
        while buf:
            out.append(buf.pop(int(u[len(out)] * len(buf))))
        return np.asarray(out, np.int64)

    def batches():
        cached = []
        for epoch in range(n_epochs):
            if epoch > 0 and is_train:
                for b in cached:
                    yield b
                continue
            order = shuffled_order()
            for s in range(0, len(order), batch_size):
                sel = order[s:s + batch_size]
                b = shards.parse(of[sel], orr[sel], num_answers, t_cap, want_image_id)
                if dense_target:
                    b["answer_target"] = densify(b)
                if is_train:
                    cached.append(b)
                yield b

    return batches()
