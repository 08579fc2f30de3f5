"""Host logic of the batch-sharded data-parallel path on CPU: world_size 2, gloo. The arithmetic claim tested
with the oracle as the stand-in for a rank's CUDA step: sum over ranks of grad(loss_r / world) over the rank's
shard == gradient of the global-batch mean loss."""
import os
import socket

import numpy as np
import pytest

torch = pytest.importorskip("torch")
import torch.multiprocessing as mp  # noqa: E402


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    from oracle import answer_model_np as O
    from vqa_transfer_externaldata_b200 import synthetic as S
    from vqa_transfer_externaldata_b200.dp import DataParallel
    dp = DataParallel(backend="gloo", bucket_bytes=4096)  # small buckets: several per tensor
    assert dp.world_size == world and dp.rank == rank
    c = S.dims(B=8, K=5, Dv=16, D=8, L=8, A=11, T=4, W=6, Vq=20)
    params, exist = S.init_params(c, seed=1, perturb=0.2)
    feats, nb = S.make_bank(c, num_images=10, seed=2, ragged_boxes=True)
    batch = S.make_batch(c, 10, seed=3)
    io, ia = S.make_answer_flags(c)
    m = O.answer_masks(c["A"], c["num_train_answer"], io, ia, exist)
    rng = np.random.default_rng(5)
    am = (rng.uniform(size=(c["B"], c["K"], c["D"])) < 0.8).astype(np.float64)
    jm = (rng.uniform(size=(c["B"], c["J"])) < 0.5).astype(np.float64)
    s, e = dp.shard(c["B"])
    assert (s, e) == (rank * 4, rank * 4 + 4)
    sub = {k: v[s:e] for k, v in batch.items()}
    _, cache = O.forward(params, feats, nb, sub, m, att_mask=am[s:e], joint_mask=jm[s:e])
    g = O.backward(cache, loss_scale=1.0 / world)
    fields = O.trainable_fields("vlmap_answer")
    flat = torch.from_numpy(np.concatenate([g[f].ravel() for f in fields]))
    dp.all_reduce_flat(flat)
    mx = dp.max_over_ranks(float(rank + 1))
    assert mx == float(world)
    dp.barrier()
    if rank == 0:
        _, cache = O.forward(params, feats, nb, batch, m, att_mask=am, joint_mask=jm)
        gg = O.backward(cache)
        ref = np.concatenate([gg[f].ravel() for f in fields])
        np.save(os.path.join(out_dir, "err.npy"), np.array([np.abs(flat.numpy() - ref).max(), np.abs(ref).max()]))
    dp.close()


def test_gradient_allreduce_world2(tmp_path):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    err, scale = np.load(os.path.join(str(tmp_path), "err.npy"))
    assert err <= 1e-12 * max(scale, 1.0), (err, scale)


def test_shard_covers_batch_without_overlap():
    from vqa_transfer_externaldata_b200.dp import DataParallel
    for world in (1, 2, 3, 8):
        for n in (0, 1, 7, 512, 513):
            seen = []
            for r in range(world):
                dp = DataParallel.__new__(DataParallel)
                dp.rank, dp.world_size = r, world
                s, e = dp.shard(n)
                seen.extend(range(s, e))
            assert seen == list(range(n))
