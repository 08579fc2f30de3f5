"""Batch-sharded data parallelism: one process per GPU, one all-reduce(sum) of the flat trainable-gradient
buffer per step (the reference has no distributed code at all: one process per GPU via
CUDA_VISIBLE_DEVICES, run.py:25-46). Each rank scales d(loss) by 1/world_size so the summed gradient is the
gradient of the global-batch mean loss. Inference shards the batch with no collective."""
import os

import torch
import torch.distributed as dist


class DataParallel:
    def __init__(self, backend=None, bucket_bytes=64 << 20):
        self.rank = int(os.environ.get("RANK", "0"))
        self.world_size = int(os.environ.get("WORLD_SIZE", "1"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        self.bucket_elems = bucket_bytes // 4
        self.owns_group = False
        self._side = None
        self._mc = None
        self._in_library = False
        if self.world_size > 1 and not dist.is_initialized():
            backend = backend or ("nccl" if torch.cuda.is_available() else "gloo")
            if backend == "nccl":
                torch.cuda.set_device(self.local_rank)
            dist.init_process_group(backend=backend)
            self.owns_group = True

    def shard(self, n_global):
        """[start, end) of this rank's contiguous slice of a global batch of n_global samples."""
        per = (n_global + self.world_size - 1) // self.world_size
        s = min(self.rank * per, n_global)
        return s, min(s + per, n_global)

    def all_reduce_flat(self, flat):
        """Sum `flat` (1-D tensor) over ranks in place, bucketed."""
        if self.world_size == 1:
            return
        n = flat.numel()
        for s in range(0, n, self.bucket_elems):
            dist.all_reduce(flat[s:min(s + self.bucket_elems, n)], op=dist.ReduceOp.SUM)

    def use_multicast_gradients(self, engine):
        """Put the engine's flat gradient buffer into symmetric memory mapped as one NVSwitch multicast object, so that
        all_reduce_gradients can run the library's in-switch all-reduce kernel (vqa_multimem_all_reduce) instead of
        NCCL's ring. Returns False (and changes nothing) when the platform has no multicast support."""
        if self.world_size == 1 or not dist.is_initialized() or dist.get_backend() != "nccl":
            return False
        if os.environ.get("VQA_DP_MULTICAST", "1") == "0":
            return False
        try:
            import torch.distributed._symmetric_memory as symm_mem
            n_grad = engine.params.n_train + engine.params.TAIL
            flags_off = (n_grad + 1023) // 1024 * 1024          # barrier counters of the in-library exchange live here
            n = flags_off + 1024
            buf = symm_mem.empty(n, dtype=torch.float32, device=engine.device)
            hdl = symm_mem.rendezvous(buf, dist.group.WORLD)
            if not hdl.has_multicast_support or not hdl.multicast_ptr:
                return False
        except Exception as e:  # noqa: BLE001 -- any failure here just means "keep using NCCL"
            if self.rank == 0:
                print(f"[dp] multicast gradients unavailable ({e!r}); using NCCL all-reduce", flush=True)
            return False
        buf.zero_()
        engine.rebind_gradients(buf)
        self._mc = (hdl, buf, n_grad)   # the tail slots are summed with the gradients
        self._in_library = False
        # in-library exchange: validated on 2 GPUs; on 8 GPUs it hit launch failures (profiles/r02_dp_n8.md), so larger jobs
        # keep the host-issued barriers of round 1 unless VQA_DP_IN_LIBRARY=1 asks for it
        if os.environ.get("VQA_DP_IN_LIBRARY", "1" if self.world_size <= 2 else "0") != "0":
            # the exchange runs inside vqa_backward (early slice under the BPTT, in-kernel barriers): nothing to do per step
            from . import lib as L
            torch.cuda.synchronize(engine.device)
            hdl.barrier(channel=0)          # every rank's flag words are zero before anybody's first backward
            torch.cuda.synchronize(engine.device)
            L.check(engine.lib.vqa_set_gradient_allreduce(engine.h, hdl.multicast_ptr, buf.data_ptr(), engine.params.n_early,
                                                          n_grad, flags_off, self.rank, self.world_size))
            self._in_library = True
        return True

    def all_reduce_gradients(self, engine):
        """One all-reduce(sum) of the trainable gradients. With early gradients enabled on the engine the slice that
        is complete before the GRU's back-propagation through time (everything but the embedding and the GRU) is
        reduced on a side stream as soon as the device reaches that point -- under the recurrent kernels -- and only
        the rest waits for the end of the backward pass."""
        g = engine.params.grad_buf   # gradients + tail slots (the embedding slice norm adds up over ranks like a gradient)
        if self.world_size == 1:
            return
        if self._mc is not None and getattr(self, "_in_library", False):
            return   # vqa_backward has already reduced them (vqa_set_gradient_allreduce)
        if self._mc is not None:
            hdl, buf, n = self._mc
            from . import lib as L
            n4 = (n + 3) // 4 * 4
            stream = torch.cuda.current_stream(g.device)
            hdl.barrier(channel=0)   # every rank's gradients are complete (stream-ordered, system scope)
            L.check(engine.lib.vqa_multimem_all_reduce(hdl.multicast_ptr, n4, self.rank, self.world_size, 0,
                                                       stream.cuda_stream))
            hdl.barrier(channel=1)   # every slice has been broadcast
            return
        n_early = engine.params.n_early if getattr(engine, "early_gradients", False) else 0
        if n_early == 0 or dist.get_backend() != "nccl":
            self.all_reduce_flat(g)
            return
        if self._side is None:
            self._side = torch.cuda.Stream(device=g.device)
        engine.stream_wait_early_gradients(self._side)
        with torch.cuda.stream(self._side):
            w1 = dist.all_reduce(g[:n_early], op=dist.ReduceOp.SUM, async_op=True)
        w2 = dist.all_reduce(g[n_early:], op=dist.ReduceOp.SUM, async_op=True) if n_early < g.numel() else None
        w1.wait()
        if w2 is not None:
            w2.wait()

    def self_check(self, device=None, per_rank=32, tol=None):
        """Hardware check of the data-parallel step on THIS job's ranks and collective: every rank runs forward +
        backward(loss_scale = 1 / world) on its shard of one global batch and the gradients are all-reduced exactly as
        in training (the in-switch multimem kernel when available, NCCL otherwise); rank 0 also runs the SAME global
        batch as one single-rank step. The two gradient sets must agree per tensor to `tol` (max-norm relative; default
        1e-5 in fp32 mode, 5e-5 in bf16 mode: the sums over the batch are formed in a different order) and be
        bit-identical across ranks. Dropout is off (keep = 1): a rank's mask is keyed by its local element index.
        Returns {precision: worst relative error}; raises AssertionError on a mismatch."""
        import numpy as np
        from . import synthetic as S
        from .engine import AnswerModelConfig, Engine
        if self.world_size == 1:
            return {}
        world, rank = self.world_size, self.rank
        device = device if device is not None else torch.device(f"cuda:{self.local_rank}")
        Bg = per_rank * world
        dims = dict(K=12, Dv=256, D=128, L=128, A=200, T=6, W=20, Vq=50)
        out = {}
        keep_mc, keep_lib = self._mc, getattr(self, "_in_library", False)
        for precision in ("fp32", "bf16"):
            limit = tol if tol is not None else (1e-5 if precision == "fp32" else 5e-5)
            cg = S.dims(B=Bg, **dims)
            params, exist = S.init_params(cg, seed=11, perturb=0.2)
            feats, nb = S.make_bank(cg, num_images=40, seed=12, ragged_boxes=True)
            batch = S.make_batch(cg, 40, seed=13)
            is_obj, is_attr = S.make_answer_flags(cg)

            def run(B, sub, scale, reduce):
                c = S.dims(B=B, **dims)
                eng = Engine(AnswerModelConfig(variant="vlmap_answer", precision=precision, keep_att=1.0, keep_joint=1.0, **c),
                             device=device)
                eng.set_feature_bank(feats, nb)
                eng.set_answer_masks(is_obj, is_attr, exist)
                eng.load_params(params)
                if reduce:
                    self._mc = None
                    self.use_multicast_gradients(eng)
                eng.stage_batch(sub)
                eng.forward(seed=1, step=1)
                eng.backward(loss_scale=scale)
                if reduce:
                    self.all_reduce_gradients(eng)
                torch.cuda.synchronize(device)
                g = {f: v.detach().clone() for f, v in eng.params.grad_views.items()}
                used_mc = ("multimem, inside vqa_backward" if self._mc is not None and self._in_library else
                           "multimem" if self._mc is not None else "nccl")
                eng.close()
                return g, used_mc

            s0, s1 = self.shard(Bg)
            g_dp, used_mc = run(s1 - s0, {k: v[s0:s1] for k, v in batch.items()}, 1.0 / world, True)
            flat = torch.cat([g_dp[f].reshape(-1) for f in sorted(g_dp)])
            ref0 = flat.clone()
            dist.broadcast(ref0, src=0)
            same = bool(torch.equal(ref0, flat))
            worst = 0.0
            if rank == 0:
                g_one, _ = run(Bg, batch, 1.0, False)
                for f in sorted(g_dp):
                    a, b = g_dp[f].double(), g_one[f].double()
                    err = float((a - b).abs().max() / max(float(b.abs().max()), 1e-30))
                    worst = max(worst, err)
            worst = self.max_over_ranks(worst)
            n_diff = self.max_over_ranks(0.0 if same else 1.0)
            assert n_diff == 0.0, f"data-parallel gradients differ between ranks ({precision})"
            assert worst <= limit, f"data-parallel step != single-rank step on the concatenated batch ({precision}): {worst:.3e} > {limit:.1e}"
            out[precision] = worst
            out["collective"] = "multimem" if used_mc else "nccl"
        self._mc, self._in_library = keep_mc, keep_lib
        return out

    def barrier(self):
        if self.world_size > 1:
            dist.barrier()

    def max_over_ranks(self, value):
        if self.world_size == 1:
            return value
        dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
        t = torch.tensor([value], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def broadcast_params(self, engine, src=0):
        if self.world_size > 1:
            dist.broadcast(engine.params.flat, src=src)
            engine.prepare_params()

    def close(self):
        if self.owns_group and dist.is_initialized():
            dist.destroy_process_group()
