"""Where does the end-to-end loop (host batches -> loss on host) lose time against the device-resident loop?"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from vqa_transfer_externaldata_b200 import synthetic as S  # noqa: E402
from vqa_transfer_externaldata_b200.model import Model, make_synthetic_config  # noqa: E402

CFG1 = dict(B=512, K=36, Dv=2048, D=1024, L=1024, A=3000, T=14, W=300, Vq=8192)


def main():
    dev = torch.device("cuda:0")
    torch.cuda.set_device(0)
    c = S.dims(**CFG1)
    n_img = 4096
    g = torch.Generator(device=dev).manual_seed(99)
    bank = torch.randn(n_img, c["K"], c["Dv"], device=dev, generator=g).abs_().mul_(0.5)
    config, _, _, _ = make_synthetic_config(CFG1, variant="vlmap_answer", precision="bf16", seed=4321, num_images=2)
    feats = {"features": bank, "num_boxes": np.full(n_img, c["K"], np.int32), "max_box_num": c["K"], "vfeat_dim": c["Dv"]}
    config.device = dev
    R = 4
    host = [S.make_batch(c, n_img, seed=1234 + 17 * r) for r in range(R)]
    keys = ("image_idx", "q_intseq", "q_intseq_len", "answer_target")
    pinned = [{k: torch.from_numpy(np.ascontiguousarray(hb[k])).pin_memory() for k in keys} for hb in host]
    devb = [{k: torch.from_numpy(np.ascontiguousarray(hb[k])).to(dev) for k in keys} for hb in host]
    model = Model(host[0], config, is_train=True, image_features=feats)
    eng = model.engine

    def loop(batches, steps, prefetch=True, read=True, bind=True, depth=1):
        pend = []
        for i in range(steps):
            nxt = batches[(i + 1) % R] if (prefetch and i + 1 < steps) else None
            eng.stage_batch(batches[i % R])
            eng.forward(seed=1, step=i, full_outputs=False)
            if bind:
                model._bind_outputs()
            if nxt is not None:
                eng.prefetch_batch(nxt)
            eng.backward()
            eng.adam_step()
            if read:
                pend.append(eng.read_scalars_async())
                if len(pend) > depth:
                    pend.pop(0).get()
        for p in pend:
            p.get()

    def timeit(name, **kw):
        loop(kw["batches"], 5, **{k: v for k, v in kw.items() if k != "batches"})
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        loop(kw["batches"], 40, **{k: v for k, v in kw.items() if k != "batches"})
        e1.record()
        torch.cuda.synchronize()
        print(f"{name:55s} {e0.elapsed_time(e1) / 40:.4f} ms/step (wall {1e3 * (time.perf_counter() - t0) / 40:.4f})", flush=True)

    def bench_like(steps):
        pending = None
        for i in range(steps):
            nxt = pinned[(i + 1) % R] if i + 1 < steps else None
            p, a, b = model.train_step(pinned[i % R], next_batch=nxt, sync=False)
            if pending is not None:
                pending.get()
            pending = p
        pending.get()

    # host time per step: how long the enqueue of one step takes when the device is never waited for
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    pend = []
    for i in range(20):
        p, _, _ = model.train_step(pinned[i % R], next_batch=pinned[(i + 1) % R], sync=False)
        pend.append(p)
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    print(f"host enqueue of Model.train_step (no waiting): {1e3 * (t1 - t0) / 20:.4f} ms/step", flush=True)
    for p in pend:
        p.get()
    import cProfile
    import pstats
    pr = cProfile.Profile()
    pr.enable()
    pend = []
    for i in range(20):
        p, _, _ = model.train_step(pinned[i % R], next_batch=pinned[(i + 1) % R], sync=False)
        pend.append(p)
    pr.disable()
    torch.cuda.synchronize()
    for p in pend:
        p.get()
    pstats.Stats(pr).sort_stats("cumulative").print_stats(18)

    for steps in (30, 30, 100):
        bench_like(3)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        bench_like(steps)
        e1.record()
        torch.cuda.synchronize()
        print(f"bench-like Model.train_step loop, {steps} steps: {e0.elapsed_time(e1) / steps:.4f} ms/step", flush=True)
    timeit("device batches, no read, no bind", batches=devb, read=False, bind=False)
    timeit("device batches, read depth 1", batches=devb, read=True, bind=False)
    timeit("device batches, read depth 2", batches=devb, read=True, bind=False, depth=2)
    timeit("pinned host batches, no read, no bind", batches=pinned, read=False, bind=False)
    timeit("pinned host batches, read depth 1, bind", batches=pinned, read=True, bind=True)
    timeit("pinned host batches, read depth 2, bind", batches=pinned, read=True, bind=True, depth=2)
    timeit("pinned host batches, no prefetch, read depth 1", batches=pinned, prefetch=False, read=True, bind=False)


if __name__ == "__main__":
    main()
