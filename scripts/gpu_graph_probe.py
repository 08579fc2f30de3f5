"""Probe: does capturing the whole cfg1 train step (forward + backward + clip/Adam + shadow refresh) in ONE CUDA
graph shorten the step?  Seed / step / Adam t are frozen in the captured graph: a timing probe only."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

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
    hb = S.make_batch(c, n_img, seed=1234)
    model = Model(hb, config, is_train=True, image_features=feats)
    eng = model.engine
    eng.stage_batch(hb)

    def step():
        eng.forward(seed=model.seed, step=3, full_outputs=False)
        model.backward()
        eng.adam_step(lr=1e-3, clip_norm=20.0)

    def timeit(fn, n=50):
        for _ in range(5):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        t1 = time.perf_counter()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n, (t1 - t0) * 1e3 / n

    ms, host = timeit(step)
    print(f"eager : {ms:.3f} ms/step device, host enqueue {host:.3f} ms/step")
    graph = torch.cuda.CUDAGraph()
    torch.cuda.synchronize()
    try:
        with torch.cuda.graph(graph, capture_error_mode="relaxed"):
            step()
    except Exception as e:  # noqa: BLE001
        print("capture failed:", repr(e))
        return 1
    torch.cuda.synchronize()
    ms, host = timeit(graph.replay)
    print(f"graph : {ms:.3f} ms/step device, host enqueue {host:.3f} ms/step")
    loss = float(eng.o_loss.item())
    print("loss after graph replays:", loss, "finite" if np.isfinite(loss) else "NOT FINITE")
    return 0


if __name__ == "__main__":
    sys.exit(main())
