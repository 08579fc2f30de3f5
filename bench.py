#!/usr/bin/env python
"""bench.py -- the reference's headline workload (BASELINE.json configs[0]/[1] shapes) on N B200s.

  python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
  python bench.py --impl reference --gpus N --steps K ...   # CPU arm (see below)
  python bench.py --mode infer --gpus N                      # BASELINE config 5: batched inference sweep, batch sharded

A "step" is one train step of vqa/trainer.py:275-287 on one synthetic batch: forward + backward of the
vlmap_answer model at B 512 x K 36 x Dv 2048 (T 14, A 3000) + gradient all-reduce (N > 1) + global-norm
clip + Adam + refresh of the bf16 weight shadows.  `value` times it with the batch already in HBM;
`e2e` times Model.train_step() fed from pinned HOST buffers (H2D of the batch + D2H of loss/report
inside the timed region).  Timing is CUDA events on the launching stream, max over ranks.

Reference arm: the reference's TF-1.6 graph cannot run here (no TensorFlow in the image), so
`--impl reference` times the oracle's PyTorch-CPU port of the same graph (oracle/answer_model_torch.py,
fp32, all host threads) on a bounded sample of the same workload; kind = "port".
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CFG1 = dict(B=512, K=36, Dv=2048, D=1024, L=1024, A=3000, T=14, W=300, Vq=8192)
WORKLOAD = "cfg1 vlmap_answer train step fwd+bwd+clip+adam, B512 x K36 x Dv2048, T14, A3000"
METRIC = "train samples/sec fwd+bwd bs512x36x2048"


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return {"hbm_gbs": d["hbm_gbs"], "bf16_tflops": d["bf16_tflops"],
                "bf16_tflops_sustained": d.get("bf16_tflops_sustained", d["bf16_tflops"]), "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


def ncu_traffic(dom):
    """DRAM bytes of the dominant section's kernel from the committed `ncu --set full` captures (None if absent: the
    cooperative + cluster recurrent kernels cannot be replayed by ncu)."""
    want = {"vproj_wgrad": "v_linear_v wgrad", "vproj_fwd": "v_linear_v forward", "attn_fwd": "attention forward",
            "attn_bwd": "attention backward"}.get(dom)
    if want is None:
        return None
    for name in ("r02_ncu_full_summary.json", "r01_ncu_full_summary_v5.json"):
        try:
            with open(os.path.join(ROOT, "profiles", name)) as f:
                for k in json.load(f)["kernels"]:
                    if k["role"].startswith(want):
                        return k["dram_traffic_bytes"]
        except (OSError, KeyError, ValueError):
            continue
    return None


_SAMPLER_SRC = r"""
import sys, time
import pynvml as n
n.nvmlInit()
h = n.nvmlDeviceGetHandleByIndex(int(sys.argv[1]))
mx = n.nvmlDeviceGetMaxClockInfo(h, n.NVML_CLOCK_SM)
reasons = getattr(n, "nvmlDeviceGetCurrentClocksEventReasons", None) or n.nvmlDeviceGetCurrentClocksThrottleReasons
while True:
    print(n.nvmlDeviceGetClockInfo(h, n.NVML_CLOCK_SM), mx, n.nvmlDeviceGetPowerUsage(h) / 1000.0, int(reasons(h)), flush=True)
    time.sleep(float(sys.argv[2]))
"""


class ClockSampler:
    """SM clock / power / throttle reasons sampled DURING the timed regions, every 25 ms, by a helper PROCESS that
    holds an NVML handle (nvidia_ml_py). Why not `nvidia-smi -lms`: it initialises NVML inside the timed region and
    stretched the end-to-end loop by ~10 %; why not a thread of this process: it competes with the enqueue loop for
    the GIL (scripts/gpu_e2e_probe.py: the same loop runs 1.04 ms/step unobserved). start() returns once the helper
    has delivered its first sample, i.e. after its NVML initialisation. nvidia-smi stays the fallback."""

    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}

    def __init__(self, index):
        self.index = index
        self.rows = []          # (sm MHz, max MHz, watts, reasons bitmask)
        self.proc = None
        self.source = None
        self._first = threading.Event()

    def start(self):
        try:
            import pynvml  # noqa: F401 -- only to know the helper can import it
            period = float(os.environ.get("VQA_BENCH_SAMPLE_MS", "25")) / 1e3
            self.proc = subprocess.Popen([sys.executable, "-c", _SAMPLER_SRC, str(self.index), str(period)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.source = "nvml"
        except Exception:  # noqa: BLE001
            q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
                 "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
                 "clocks_event_reasons.sw_power_cap")
            try:
                self.proc = subprocess.Popen(
                    ["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "100", "-i",
                     str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
                self.source = "nvidia-smi"
            except OSError:
                self.proc = None
                return
        self.thread = threading.Thread(target=self._read, daemon=True)
        self.thread.start()
        self._first.wait(timeout=10.0)

    def _read(self):
        bits = [0x8, 0x40, 0x20, 0x4]
        for line in self.proc.stdout:
            try:
                if self.source == "nvml":
                    f = line.split()
                    self.rows.append((float(f[0]), float(f[1]), float(f[2]), int(f[3])))
                else:
                    f = [x.strip() for x in line.strip().split(",")]
                    mask = sum(bit for bit, v in zip(bits, f[3:7]) if v.lower().startswith("active"))
                    self.rows.append((float(f[0]), float(f[1]), float(f[2]), mask))
                self._first.set()
            except (ValueError, IndexError):
                continue

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no NVML / nvidia-smi"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:  # noqa: BLE001
            self.proc.kill()
        sm = [r[0] for r in self.rows]
        power = [r[2] for r in self.rows]
        reasons = set()
        for r in self.rows:
            for bit, name in self.REASONS.items():
                if r[3] & bit:
                    reasons.add(name)
        # samples under load only: the SM clock idles low between regions
        load = [s_ for s_, p_ in zip(sm, power) if p_ > 300.0] or sm
        return {"sm_mhz": float(np.median(load)) if load else None,
                "sm_max_mhz": max(r[1] for r in self.rows) if self.rows else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons),
                "source": self.source}


# ------------------------------------------------------------------------------------------------------
# CPU arm: the oracle's torch port of the same graph
# ------------------------------------------------------------------------------------------------------
def cpu_port_rate(sample_b, steps, warmup, seed=0, budget_s=None):
    """samples/sec of fwd+bwd of the torch-CPU port at cfg1 layer sizes on a batch of sample_b. budget_s: keep timing
    steps (at least 3) until that many seconds of CPU work have been spent, instead of a fixed step count."""
    import torch
    from oracle import answer_model_torch as OT
    from vqa_transfer_externaldata_b200 import synthetic as S
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    c = S.dims(**dict(CFG1, B=sample_b))
    params, _ = S.init_params(c, seed=4321)
    feats, nb = S.make_bank(c, num_images=sample_b, seed=99)
    batch = S.make_batch(c, sample_b, seed=1234)
    tp = {k: torch.tensor(v, requires_grad=True) for k, v in params.items()}
    for k in ("pl_w", "pl_b", "pl_gamma", "pl_beta", "ql_w", "ql_b", "ql_gamma", "ql_beta", "joint_w",
              "joint_b", "joint_gamma", "joint_beta", "ans_w", "ans_b"):
        tp[k].requires_grad_(False)  # frozen transfer head (vqa/model_vlmap_answer.py:81-89)
    tb = {k: torch.tensor(v) for k, v in batch.items()}
    tf, tn = torch.tensor(feats), torch.tensor(nb)
    tm = torch.tensor((np.arange(c["A"]) < c["num_train_answer"]).astype(np.float32))
    g = torch.Generator().manual_seed(seed)
    times = []
    i = -1
    while True:
        i += 1
        if budget_s is None and i >= warmup + steps:
            break
        if budget_s is not None and len(times) >= 3 and float(np.sum(times)) >= budget_s:
            break
        am = (torch.rand(sample_b, c["K"], c["D"], generator=g) < 0.8).float()
        jm = (torch.rand(sample_b, c["J"], generator=g) < 0.5).float()
        t0 = time.perf_counter()
        out = OT.forward(tp, tf, tn, tb, tm, att_mask=am, joint_mask=jm)
        out["loss"].backward()
        for p in tp.values():
            p.grad = None
        if i >= warmup:
            times.append(time.perf_counter() - t0)
    return sample_b / float(np.median(times)), threads, float(np.sum(times))


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    sample_b = CFG1["B"]   # the SAME batch as our arm (512): ~0.3 s per step on 16 threads
    rate, threads, busy = cpu_port_rate(sample_b, max(1, args.steps), max(1, min(args.warmup, 2)))
    line = {
        "impl": "reference", "metric": METRIC, "value": rate, "unit": "samples/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * sample_b / rate,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "per_gpu_batch": sample_b, "global_batch": sample_b,
                   "note": "CPU arm: the oracle's torch-CPU port of the same graph at the same batch (B 512), all host threads"},
        "cpu_baseline": {"value": rate, "unit": "samples/s", "cores": threads, "kind": "port",
                         "sample": f"fwd+bwd of oracle/answer_model_torch.py at cfg1 layer sizes, batch {sample_b}, "
                                   f"{args.steps} timed steps ({busy:.1f} s CPU wall)"},
        "e2e": {"value": rate, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    _emit(line)
    return 0



def phase_work(c, precision):
    """Algorithmic work of every section of the step's critical path:
    name -> (bound, FLOPs or bytes per step, kernel that dominates the section)."""
    B, K, Dv, D, L, J, A, T, W = (c[k] for k in ("B", "K", "Dv", "D", "L", "J", "A", "T", "W"))
    zb = vb = 2 if precision == "bf16" else 4
    att_fwd = B * (K * D * zb + K * Dv * vb + D * 4 + Dv * 4 + Dv * vb + K * 4)
    att_bwd = B * (K * Dv * vb + K * D * zb + Dv * 4 + D * 4 + K * 4 + K * D * vb + D * 4 + (4 * D + 8) * 4)
    pair = "gemm_pair_kernel" if precision == "bf16" else "gemm_bf16_tcgen05_kernel (hi/lo planes)"
    small = "gemm_bf16_tcgen05_kernel (M = 512 head GEMMs) + row_ln_relu kernels"
    return {
        "gather": ("hbm", 2.0 * B * K * Dv * vb, "gather_features_bf16_kernel"),
        "vproj_fwd": ("tensor", 2.0 * B * K * Dv * D, pair),
        "gru_fwd": ("tensor", 2.0 * B * L * 3 * L * T, "gru_pair_kernel<0> (recurrence, h-part: 2 B L 3L T)"),
        "qheads_fwd": ("tensor", 2.0 * B * (L * L + L * D), small),
        "attn_fwd": ("hbm", float(att_fwd), "attn_fwd_pipe_kernel"),
        "head_fwd": ("tensor", 2.0 * B * (Dv * L + L * J + J * A), small),
        "loss": ("hbm", 2.0 * B * A * 4, "bce_metrics_kernel"),
        "head_bwd": ("tensor", 2.0 * B * (A * J + J * L + L * Dv + L * L), small),
        "attn_bwd": ("hbm", float(att_bwd), "attn_bwd_kernel"),
        "qv_bwd": ("tensor", 2.0 * B * D * L, small),
        "vproj_wgrad": ("tensor", 2.0 * B * K * Dv * D, pair),
        "gru_bwd": ("tensor", 2.0 * B * L * 3 * L * T, "gru_pair_kernel<1> (BPTT, 2 B L 3L T)"),
        "gru_wgrad": ("tensor", 2.0 * T * B * 3 * L * (W + L) + 2.0 * T * B * W * 3 * L + 2.0 * B * K * Dv * D,
                      pair + " x 7 on forked streams (4 GRU weight gradients + dE + dWv)"),
        "embed_bwd": ("hbm", 0.0, "embed_scatter_add"),
    }, att_fwd, att_bwd

# ------------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------------
def measure_fp32_mode(bank, c, n_img, dev, peaks, steps=10, warmup=3):
    """The same train step in the reference-precision mode (north_star: fp32 mode within 1e-4), >= 10 timed steps."""
    import torch
    from vqa_transfer_externaldata_b200 import synthetic as S
    from vqa_transfer_externaldata_b200.model import Model, make_synthetic_config
    config, _, _, _ = make_synthetic_config(CFG1, variant="vlmap_answer", precision="fp32", seed=4321, num_images=2)
    config.device = dev
    feats = {"features": bank, "num_boxes": np.full(n_img, c["K"], np.int32), "max_box_num": c["K"], "vfeat_dim": c["Dv"]}
    hb = [S.make_batch(c, n_img, seed=4242 + r) for r in range(2)]
    m = Model(hb[0], config, is_train=True, image_features=feats)
    e = m.engine
    db = [{k: torch.from_numpy(np.ascontiguousarray(b[k])).to(dev) for k in ("image_idx", "q_intseq", "q_intseq_len", "answer_target")}
          for b in hb]

    def step(i):
        e.stage_batch(db[i % 2])
        e.forward(seed=m.seed, step=m.global_step, full_outputs=False, defer_outputs=True)
        m.backward()
        e.adam_step(lr=1e-3, clip_norm=20.0)
        m.global_step += 1

    for i in range(warmup):
        step(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        step(i)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    loss, _ = e.read_scalars()
    t_roof = (3 * 356.1e9 / (peaks["bf16_tflops_sustained"] * 1e12) + (543.4e6 + 18.4e6) / (peaks["hbm_gbs"] * 1e9)) * 1e3
    out = {"ms_per_step": ms, "samples_per_s": c["B"] / (ms * 1e-3), "steps": steps, "warmup": warmup,
           "step_roofline_frac": t_roof / ms, "t_roof_ms": t_roof, "loss_finite": bool(np.isfinite(loss)),
           "model": "3 x 356.1 GF / bf16 sustained (x = hi + lo bf16 planes, 3 tcgen05 MMAs per k-step, fp32 TMEM accumulation) "
                    "+ 561.8 MB / HBM"}
    e.close()
    del m, e
    torch.cuda.empty_cache()
    return out


INFER_DIMS = dict(K=100, Dv=2048, D=1024, L=1024, A=3000, T=14, W=300, Vq=8192)


def infer_sweep(dp, dev, precision, sizes, steps, warmup):
    """BASELINE config 5 (vqa/evaler.py:118-123): forward-only sweep over GLOBAL batch sizes, K = 100 padded boxes with
    10..100 valid per image, the batch sharded over the ranks with NO collective on the path. Returns (rows, launches)."""
    import torch
    from vqa_transfer_externaldata_b200 import synthetic as S
    from vqa_transfer_externaldata_b200.engine import AnswerModelConfig, Engine
    rank, world = dp.rank, dp.world_size
    n_img = 2048
    g = torch.Generator(device=dev).manual_seed(99 + rank)
    K = INFER_DIMS["K"]
    bank = torch.randn(n_img, K, INFER_DIMS["Dv"], device=dev, generator=g).abs_().mul_(0.5)
    rng = np.random.default_rng(5 + rank)
    nb = rng.integers(10, K + 1, size=n_img).astype(np.int32)
    bank *= (torch.arange(K, device=dev)[None, :] < torch.from_numpy(nb).to(dev)[:, None])[:, :, None]
    sweep = []
    launches = 0
    keys = ("image_idx", "q_intseq", "q_intseq_len", "answer_target")
    for Bg in sizes:
        s0, s1 = dp.shard(Bg)
        Bl = s1 - s0
        c = S.dims(B=max(Bl, 1), **INFER_DIMS)
        eng = Engine(AnswerModelConfig(variant="vlmap_answer", precision=precision, **c), device=dev)
        eng.set_feature_bank(bank, nb)
        params, exist = S.init_params(c, seed=4321, variant="vlmap_answer")
        is_obj, is_attr = S.make_answer_flags(c)
        eng.set_answer_masks(is_obj, is_attr, exist)
        eng.load_params(params)
        hb = [S.make_batch(c, n_img, seed=1234 + 17 * r + 1000 * rank) for r in range(3)]
        db = [{k: torch.from_numpy(np.ascontiguousarray(b[k])).to(dev) for k in keys} for b in hb]
        # host batches as the native reader delivers them: the soft-score target as sparse triples (densified on the device)
        pinned = []
        for b in hb:
            pb = {k: torch.from_numpy(np.ascontiguousarray(b[k])).pin_memory() for k in keys[:3]}
            rows, ids = np.nonzero(b["answer_target"])
            pb["answer_sparse"] = (rows.astype(np.int32), ids.astype(np.int32), b["answer_target"][rows, ids].astype(np.float32))
            pinned.append(pb)
        h2d_sparse = 12 * max(len(pb["answer_sparse"][0]) for pb in pinned)
        host_out = torch.zeros(c["B"], dtype=torch.int32).pin_memory()

        def step_dev(i):
            eng.stage_batch(db[i % 3])
            eng.forward(seed=777 + 7919 * rank, step=i, full_outputs=True)

        def step_e2e(i):   # host batch in, predictions back on the host
            eng.stage_batch(pinned[i % 3])
            eng.forward(seed=777 + 7919 * rank, step=i, full_outputs=True)
            host_out[:Bl].copy_(eng.o_pred[:Bl], non_blocking=True)

        def timed(fn, n):
            dp.barrier()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            n0 = eng.launch_count()
            e0.record()
            for i in range(n):
                fn(i)
            e1.record()
            torch.cuda.synchronize()
            dp.barrier()
            return dp.max_over_ranks(e0.elapsed_time(e1)) / n, eng.launch_count() - n0

        for i in range(max(3, warmup)):
            step_dev(i)
        ms, n_l = timed(step_dev, steps)
        launches += n_l
        for i in range(2):
            step_e2e(i)
        ms_e2e, _ = timed(step_e2e, steps)
        h2d = Bl * 8 + Bl * c["T"] * 4 + Bl * 4 + h2d_sparse
        ok = bool(torch.isfinite(eng.outputs()["att_score"]).all().item())
        sweep.append({"global_batch": Bg, "per_gpu_batch": Bl, "ms_per_step": ms, "samples_per_s": Bg / (ms * 1e-3),
                      "e2e_ms_per_step": ms_e2e, "e2e_samples_per_s": Bg / (ms_e2e * 1e-3),
                      "h2d_bytes_per_step": h2d * world, "d2h_bytes_per_step": 4 * Bl * world,
                      "roofline_frac": 0.59e-3 * Bl / ms, "outputs_finite": ok,
                      "gru_kernels": {1: "pair", 2: "single-CTA"}.get(int(eng.lib.vqa_gru_kernel_path()) & 3, "?")})
        eng.close()
        del eng
        torch.cuda.empty_cache()
    return sweep, launches


def run_infer(args):
    """`--mode infer`: the BASELINE config 5 sweep as its own JSON line."""
    import torch
    from vqa_transfer_externaldata_b200.dp import DataParallel
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- this repo has no CPU path")
    dp = DataParallel()
    rank, world = dp.rank, dp.world_size
    torch.cuda.set_device(dp.local_rank)
    dev = torch.device(f"cuda:{dp.local_rank}")
    sizes = [int(x) for x in args.infer_batches.split(",")]
    sampler = ClockSampler(dp.local_rank)
    if rank == 0:
        sampler.start()
    sweep, launches = infer_sweep(dp, dev, args.precision, sizes, args.steps, args.warmup)
    clocks = sampler.stop() if rank == 0 else None
    if rank == 0:
        best = max(sweep, key=lambda r: r["samples_per_s"])
        line = {"metric": "inference samples/sec fwd, K100 masked, batch sharded over GPUs (BASELINE config 5)",
                "value": best["samples_per_s"], "unit": "samples/s", "n_gpus": world, "steps": args.steps,
                "warmup": max(3, args.warmup), "ms_per_step": best["ms_per_step"], "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None, "dtype": "bf16" if args.precision == "bf16" else "f32",
                "data": "synthetic",
                "config": {"workload": "cfg5 vlmap_answer forward (logits, attention, pred, report), K 100 padded boxes with "
                                       "10..100 valid, T14, A3000; global batch sharded over the ranks, no collective",
                           "global_batch": best["global_batch"], "parallelism": f"shard{world}", "precision": args.precision,
                           "l2_policy": "inputs larger than L2: 2048-image x 100-box feature bank (1.68 GB fp32 + 0.84 GB bf16 copy) "
                                        "indexed at random", "roofline_per_sample_us": 0.59},
                "e2e": {"value": best["e2e_samples_per_s"], "unit": "samples/s", "ms_per_step": best["e2e_ms_per_step"],
                        "h2d_bytes_per_step": best["h2d_bytes_per_step"], "d2h_bytes_per_step": best["d2h_bytes_per_step"]},
                "gpu_launches": int(launches), "clocks": clocks, "sweep": sweep}
        _emit(line)
    dp.close()
    return 0


def memft_cpu_rate(n_images=32, steps=2, budget_s=25.0):
    """CPU baseline of BASELINE config 4: the oracle's torch-autograd twin of the pre-training graph
    (oracle/memft_torch.py, fp32, all host threads) forward + backward on a bounded sample (n_images images of the cfg4
    shapes). The reference's TF-1.6 graph cannot run in this image."""
    import torch
    from oracle import memft_np as M
    from oracle import memft_torch as MT
    from vqa_transfer_externaldata_b200 import memft as F
    dims = dict(F.CFG4, B=n_images)
    cfg = F.make_config(dims)
    p = {k: torch.tensor(v, dtype=torch.float32, requires_grad=True) for k, v in F.xavier_params(cfg).items()}
    hb = F.synthetic_batch(dims, seed=7)
    tb = {k: torch.tensor(v, dtype=torch.float32 if np.asarray(v).dtype.kind == "f" else torch.int64) for k, v in hb.items()}
    tm = {k: torch.tensor(v, dtype=torch.float32) for k, v in M.make_masks(dims, seed=3).items()}
    times = []
    t_start = time.perf_counter()
    for i in range(steps + 1):
        t0 = time.perf_counter()
        loss, _ = MT.forward(p, tb, tm)
        loss.backward()
        for v in p.values():
            v.grad = None
        if i:
            times.append(time.perf_counter() - t0)
        if time.perf_counter() - t_start > budget_s and times:
            break
    return n_images / float(np.median(times)), torch.get_num_threads(), float(np.sum(times))


def measure_memft(dp, dev, precision, peaks, steps=5, warmup=3):
    """A short data-parallel run of the cfg4 train step on the ranks of this job (needs the engine of the headline run
    closed: the two workspaces do not fit side by side with the 4096-image bank). Returns a dict on every rank."""
    import torch
    from vqa_transfer_externaldata_b200 import memft as F
    torch.cuda.empty_cache()
    cfg = F.make_config(F.CFG4, precision=precision)
    model = F.Model(F.synthetic_batch(F.CFG4, seed=100 + 1000 * dp.rank), cfg, is_train=True, params=F.xavier_params(cfg), seed=777)
    for _ in range(warmup):
        model.train_step(sync=False)
    dp.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        model.train_step(sync=False)
    e1.record()
    torch.cuda.synchronize()
    dp.barrier()
    ms = dp.max_over_ranks(e0.elapsed_time(e1)) / steps
    loss, _ = model.fetch()
    flops = F.gemm_flops_per_step(cfg)
    model.close()
    del model
    torch.cuda.empty_cache()
    B = F.CFG4["B"]
    return {"config": "cfg4 vlmap_memft bf_or_wordset_withatt_sp train step fwd+bwd+clip+adam, B512 per GPU x (5+5) entries, K36 x Dv2048, "
                      "5120 blank sequences T10, A4000; batch-sharded DP, NCCL all-reduce of the flat gradient",
            "ms_per_step": ms, "images_per_s": B * dp.world_size / ms * 1e3, "steps": steps, "warmup": warmup,
            "tflops_per_gpu": flops / 1e9 / ms, "roofline_frac": flops / (peaks["bf16_tflops_sustained"] * 1e12) * 1e3 / ms,
            "loss_finite": bool(np.isfinite(loss))}


def run_memft(args):
    """`--mode memft`: BASELINE config 4 -- one train step of the vlmap pre-training graph (bs 512 per GPU, 5 + 5 entries
    per image, K 36, A 4000, blanks <= 10 tokens), batch-sharded data parallel over N GPUs (NCCL all-reduce of the flat
    gradient)."""
    import torch
    from vqa_transfer_externaldata_b200 import lib as L
    from vqa_transfer_externaldata_b200 import memft as F
    from vqa_transfer_externaldata_b200.dp import DataParallel
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- this repo has no CPU path")
    dp = DataParallel()
    rank, world = dp.rank, dp.world_size
    torch.cuda.set_device(dp.local_rank)
    dev = torch.device(f"cuda:{dp.local_rank}")
    peaks = load_peaks()
    cfg = F.make_config(F.CFG4, precision=args.precision)
    R = 2
    host = [F.synthetic_batch(F.CFG4, seed=100 + 17 * r + 1000 * rank) for r in range(R)]
    model = F.Model(host[0], cfg, is_train=True, params=F.xavier_params(cfg), seed=777)
    lib = L.load()
    # `value`: the batch resident in HBM; `e2e`: features resident (as the answer model's bank is), everything else --
    # boxes, blanks, lengths, fills, counts, wordset ids -- uploaded from host memory every step, loss read back
    feats = [{k: torch.from_numpy(hb[k]).to(dev) for k in ("image_ft", "spatial_ft")} for hb in host]
    e2e_batches = [dict(hb, **f) for hb, f in zip(host, feats)]
    h2d = sum(int(np.asarray(v).nbytes) for k, v in host[0].items() if k not in ("image_ft", "spatial_ft"))
    W = max(3, args.warmup)

    def timed(fn, steps):
        dp.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n0 = lib.vqa_launch_count()
        e0.record()
        for i in range(steps):
            fn(i)
        e1.record()
        torch.cuda.synchronize()
        dp.barrier()
        return dp.max_over_ranks(e0.elapsed_time(e1)), lib.vqa_launch_count() - n0

    def step_resident(i):
        model.train_step(sync=False)

    def step_e2e(i):
        model.set_batch(e2e_batches[i % R])
        model.train_step(sync=True)

    timed(step_resident, W)
    sampler = ClockSampler(dp.local_rank)
    if rank == 0:
        sampler.start()
    ms, launches = timed(step_resident, args.steps)
    timed(step_e2e, 2)
    ms_e2e, _ = timed(step_e2e, args.steps)
    clocks = sampler.stop() if rank == 0 else None
    loss, report = model.fetch()
    ms /= args.steps
    ms_e2e /= args.steps
    flops = F.gemm_flops_per_step(cfg)
    cpu = None
    if rank == 0 and not args.no_cpu_baseline:
        v, cores, secs = memft_cpu_rate()
        cpu = {"value": v, "unit": "images/s", "cores": cores, "kind": "port",
               "sample": f"fwd+bwd of the oracle's torch twin of the pre-training graph at cfg4 layer sizes, 32 images per step, {secs:.1f} s of timed CPU work"}
    if rank == 0:
        B = F.CFG4["B"]
        t_roof = flops / (peaks["bf16_tflops_sustained"] * 1e12) * 1e3
        line = {"metric": "vlmap_memft pretrain images/sec fwd+bwd bs512 (5 obj + 5 attr entries, K36, A4000, blanks<=10) (BASELINE config 4)",
                "value": B * world / ms * 1e3, "unit": "images/s", "n_gpus": world, "steps": args.steps, "warmup": W,
                "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "bf16" if args.precision == "bf16" else "f32", "data": "synthetic",
                "config": {"workload": "cfg4 vlmap_memft bf_or_wordset_withatt_sp train step fwd+bwd+clip+adam, B512 x (5+5) entries, "
                                       "K36 x Dv2048, 5120 blank sequences T10, A4000", "per_gpu_batch": B, "global_batch": B * world,
                           "precision": args.precision, "parallelism": f"dp{world}",
                           "gradient_collective": "NCCL all-reduce of the flat gradient buffer" if world > 1 else "none (1 GPU)",
                           "l2_policy": "inputs larger than L2: 151 MB of image features + 164 MB of logits per step"},
                "e2e": {"value": B * world / ms_e2e * 1e3, "unit": "images/s", "ms_per_step": ms_e2e,
                        "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 52,
                        "note": "image / box features resident in HBM; ids, boxes, blanks, counts uploaded and validated on the host every step"},
                "gpu_launches": int(launches), "clocks": clocks, "loss": loss, "loss_finite": bool(np.isfinite(loss)),
                "roofline": {"bound": "tensor", "kernel": "whole step (dense contractions of the graph)", "achieved": flops / 1e9 / ms,
                             "peak": peaks["bf16_tflops_sustained"], "unit": "TFLOP/s", "frac": t_roof / ms, "traffic": None,
                             "model": f"{flops / 1e12:.2f} TFLOP per step / bf16 sustained"},
                "cpu_baseline": cpu}
        _emit(line)
    model.close()
    dp.close()
    return 0


def run_ours(args):
    import torch
    from vqa_transfer_externaldata_b200 import lib as L
    from vqa_transfer_externaldata_b200 import synthetic as S
    from vqa_transfer_externaldata_b200.dp import DataParallel
    from vqa_transfer_externaldata_b200.model import Model, make_synthetic_config

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- this repo has no CPU path (use --impl reference for the CPU arm)")
    dp = DataParallel()
    rank, world = dp.rank, dp.world_size
    torch.cuda.set_device(dp.local_rank)
    dev = torch.device(f"cuda:{dp.local_rank}")
    peaks = load_peaks()
    c = S.dims(**CFG1)
    B = c["B"]

    # synthetic bank generated ON the device (1.2 GB, larger than the 126 MB L2), params / batches in NumPy
    n_img = args.bank_images
    g = torch.Generator(device=dev).manual_seed(99 + rank)
    bank = torch.randn(n_img, c["K"], c["Dv"], device=dev, generator=g).abs_().mul_(0.5)
    config, image_features, _, _ = make_synthetic_config(CFG1, variant="vlmap_answer", precision=args.precision,
                                                         seed=4321, num_images=2)
    image_features = {"features": bank, "num_boxes": np.full(n_img, c["K"], np.int32),
                      "max_box_num": c["K"], "vfeat_dim": c["Dv"]}
    config.device = dev
    R = 4
    host_batches = [S.make_batch(c, n_img, seed=1234 + 17 * r + 1000 * rank) for r in range(R)]
    model = Model(host_batches[0], config, is_train=True, image_features=image_features)
    model.attach_data_parallel(dp if world > 1 else None)
    eng = model.engine
    dp.broadcast_params(eng)
    lib = eng.lib

    # device-resident copies of the R batches for the `value` loop (inputs already in HBM when the timed region starts)
    dev_batches = [{k: torch.from_numpy(np.ascontiguousarray(hb[k])).to(dev)
                    for k in ("image_idx", "q_intseq", "q_intseq_len", "answer_target")} for hb in host_batches]

    def step_resident(i):
        # the same software pipeline as Model.train_step: batch i is staged (adopted if step i-1 prefetched it), and
        # batch i+1's buffers + feature gather are issued right after this step's forward
        eng.stage_batch(dev_batches[i % R])
        rk = dp.rank if world > 1 else 0
        eng.forward(seed=model.seed + 7919 * rk, step=model.global_step, full_outputs=False, defer_outputs=True)
        eng.prefetch_batch(dev_batches[(i + 1) % R])
        model.backward()
        eng.adam_step(lr=1e-3, clip_norm=20.0, pipelined_tail=True)   # as Model.train_step does
        model.global_step += 1

    def timed(fn, steps):
        dp.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n0 = eng.launch_count()
        e0.record()
        for i in range(steps):
            fn(i)
        e1.record()
        torch.cuda.synchronize()
        dp.barrier()
        ms = dp.max_over_ranks(e0.elapsed_time(e1))
        return ms, eng.launch_count() - n0

    if world > 1:
        # data-parallel step == single-rank step on the concatenated batch, on THIS job's ranks and collective, before
        # anything is timed (raises on mismatch)
        dp_check = dp.self_check(device=dev)
    else:
        dp_check = None
    collective = ("none (1 GPU)" if world == 1 else
                  ("multimem.ld_reduce/st in-switch all-reduce (csrc/collective.cu) launched INSIDE vqa_backward: early slice under the "
                   "BPTT, in-kernel barriers; NCCL for rendezvous + broadcast only" if dp._in_library else
                   "multimem.ld_reduce/st in-switch all-reduce (csrc/collective.cu), NCCL for rendezvous + broadcast only")
                  if dp._mc is not None else "NCCL all-reduce")

    W = max(3, args.warmup)
    for i in range(W):
        step_resident(i)
    torch.cuda.synchronize()

    sampler = ClockSampler(dp.local_rank)
    if rank == 0:
        sampler.start()
    ms_value, launches = timed(step_resident, args.steps)

    # end to end through the public API (Model.train_step): every step uploads ITS batch from pinned host
    # memory (on the copy stream, overlapping the previous step's kernels) and its loss/report are read back
    # by the host one step later (asynchronous dispatch), all inside the timed region.
    # The soft-score target travels the way the native reader delivers it (input_native.create / csrc/input_host.cu): as
    # (row, answer id, score) triples that vqa_densify_targets scatters on the device -- ~70 KB per step instead of the
    # 6.2 MB dense [B, A] matrix (--e2e-dense uploads the dense matrix as the pure-Python reader's batches do).
    def host_batch(hb):
        out = {k: torch.from_numpy(np.ascontiguousarray(hb[k])).pin_memory() for k in ("image_idx", "q_intseq", "q_intseq_len")}
        if args.e2e_dense:
            out["answer_target"] = torch.from_numpy(np.ascontiguousarray(hb["answer_target"])).pin_memory()
        else:
            rows, ids = np.nonzero(hb["answer_target"])
            out["answer_sparse"] = (rows.astype(np.int32), ids.astype(np.int32), hb["answer_target"][rows, ids].astype(np.float32))
        return out

    pinned = [host_batch(hb) for hb in host_batches]
    h2d = d2h = 0
    losses = []

    def run_e2e(steps):
        # asynchronous dispatch, two steps deep: the host reads the loss of step i - 2 after enqueuing step i, so a
        # hiccup of the host (this loop is ~0.6 ms of Python + launches per 1.04 ms step) does not starve the device
        nonlocal h2d, d2h
        pending = []
        for i in range(steps):
            nxt = pinned[(i + 1) % R]   # steady-state pipeline: every step uploads the batch of the step after it
            p, a, b = model.train_step(pinned[i % R], next_batch=nxt, sync=False)
            h2d, d2h = a, b
            pending.append(p)
            if len(pending) > 2:
                losses.append(pending.pop(0).get()[0])
        for p in pending:
            losses.append(p.get()[0])

    # warm-up: a multiple of R steps (>= --warmup), so that the last one prefetches pinned[0], the first batch of the timed run
    run_e2e(R * ((max(R, args.warmup) + R - 1) // R))
    import gc
    gc.collect()
    gc.disable()   # (a collection inside 20 timed steps of ~1 ms each is a host stall the pipeline cannot hide)
    dp.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    run_e2e(args.steps)
    e1.record()
    torch.cuda.synchronize()
    dp.barrier()
    ms_e2e = dp.max_over_ranks(e0.elapsed_time(e1))
    gc.enable()
    assert all(np.isfinite(losses)), "non-finite loss in the e2e loop"
    clocks = sampler.stop() if rank == 0 else None

    # per-phase device times inside the real step (CUDA events recorded by the library on this stream)
    L.check(lib.vqa_profile_enable(eng.h, 1))
    keep_prefetch, eng.prefetch_features = eng.prefetch_features, False   # isolated phase times: no concurrent gather
    import ctypes as C
    acc = np.zeros(L.NUM_PHASES)
    PROF_STEPS = 5
    for i in range(PROF_STEPS):
        step_resident(i)
        buf = (C.c_float * L.NUM_PHASES)()
        L.check(lib.vqa_profile_read(eng.h, buf))
        acc += np.array(list(buf))
    phase_ms = {lib.vqa_phase_name(i).decode(): float(acc[i] / PROF_STEPS) for i in range(L.NUM_PHASES)}
    # the same events with the forked branches kept: sections of the main stream's critical path in the REAL step
    eng.prefetch_features = keep_prefetch
    L.check(lib.vqa_profile_enable(eng.h, 2))
    acc2 = np.zeros(L.NUM_PHASES)
    for i in range(PROF_STEPS):
        step_resident(i)
        buf = (C.c_float * L.NUM_PHASES)()
        L.check(lib.vqa_profile_read(eng.h, buf))
        acc2 += np.array(list(buf))
    L.check(lib.vqa_profile_enable(eng.h, 0))
    critical_ms = {lib.vqa_phase_name(i).decode(): float(acc2[i] / PROF_STEPS) for i in range(L.NUM_PHASES)}

    # optimizer and all-reduce sections of the same step (events on the launching stream)
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    opt_ms = ar_ms = 0.0
    for i in range(PROF_STEPS):
        eng.stage_batch(dev_batches[i % R])
        rk = dp.rank if world > 1 else 0
        eng.forward(seed=model.seed + 7919 * rk, step=model.global_step, full_outputs=False, defer_outputs=True)
        eng.prefetch_batch(dev_batches[(i + 1) % R])
        eng.backward(loss_scale=1.0 / world)
        evs[0].record()
        if world > 1:
            dp.all_reduce_gradients(eng)
        evs[1].record()
        eng.adam_step(lr=1e-3, clip_norm=20.0)
        evs[2].record()
        model.global_step += 1
        torch.cuda.synchronize()
        ar_ms += evs[0].elapsed_time(evs[1]) / PROF_STEPS
        opt_ms += evs[1].elapsed_time(evs[2]) / PROF_STEPS
    phase_ms["allreduce"] = critical_ms["allreduce"] = dp.max_over_ranks(ar_ms) if world > 1 else 0.0
    phase_ms["optimizer"] = critical_ms["optimizer"] = opt_ms

    samples = world * B * args.steps
    value = samples / (ms_value * 1e-3)
    e2e_value = samples / (ms_e2e * 1e-3)

    # rooflines. `roofline` names the section with the LARGEST share of the step's critical path; its work is the
    # algorithmic FLOPs / bytes of SURVEY 8d (phase_work), its time the section's device time inside the real step.
    work, att_fwd_bytes, att_bwd_bytes = phase_work(c, args.precision)
    peak_tc = peaks["bf16_tflops_sustained"] / (3.0 if args.precision == "fp32" else 1.0)
    sections = {}
    for name, (bound, amount, kern) in work.items():
        ms = critical_ms.get(name, 0.0)
        note = None
        if name == "vproj_fwd" and phase_ms.get(name, 0.0) > 0.0:
            # in the real step the first rows of this product run under the recurrent kernel (csrc/model.cu:
            # vproj_split_plan), so its critical-path window holds only part of the FLOPs: rate it on the serial timing
            ms, note = phase_ms[name], "whole product timed serially (phase_ms); in the step its first rows run under the recurrent kernel"
        if ms <= 0.0 or amount <= 0.0:
            continue
        if bound == "tensor":
            ach, peak, unit = amount / (ms * 1e-3) / 1e12, peak_tc, "TFLOP/s"
        else:
            ach, peak, unit = amount / (ms * 1e-3) / 1e9, peaks["hbm_gbs"], "GB/s"
        sections[name] = {"bound": bound, "kernel": kern, "ms": ms, "share_of_step": ms / (ms_value / args.steps),
                          "achieved": ach, "peak": peak, "unit": unit, "frac": ach / peak}
        if note:
            sections[name]["note"] = note
    dom = max(sections, key=lambda k: sections[k]["ms"])
    d = sections[dom]
    roofline = {"bound": d["bound"], "kernel": d["kernel"], "section": dom, "achieved": d["achieved"], "peak": d["peak"],
                "unit": d["unit"], "frac": d["frac"], "share_of_step": d["share_of_step"], "ms": d["ms"],
                "traffic": ncu_traffic(dom) if args.precision == "bf16" else None,
                "traffic_source": "dram__bytes_read.sum + dram__bytes_write.sum of that launch from the committed ncu --set full "
                                  "capture under profiles/ (null for the cooperative + cluster recurrent kernels, which ncu "
                                  "cannot replay: their evidence is the in-kernel globaltimer phase trace under profiles/)",
                "peak_source": f"{peaks['source']}: bf16 sustained (kernel timed inside a long step)"
                               + (" / 3 (hi + lo planes: 3 MMAs per k-step)" if args.precision == "fp32" else "")
                               if d["bound"] == "tensor" else f"{peaks['source']} HBM copy bandwidth",
                "how_chosen": "largest critical_path_ms section of the step"}
    attn = {
        "fwd": {"ms": phase_ms["attn_fwd"], "bytes": att_fwd_bytes,
                "GBps": att_fwd_bytes / (phase_ms["attn_fwd"] * 1e-3) / 1e9},
        "bwd": {"ms": phase_ms["attn_bwd"], "bytes": att_bwd_bytes,
                "GBps": att_bwd_bytes / (phase_ms["attn_bwd"] * 1e-3) / 1e9},
        "peak_GBps": peaks["hbm_gbs"],
    }
    attn["fwd"]["frac"] = attn["fwd"]["GBps"] / peaks["hbm_gbs"]
    attn["bwd"]["frac"] = attn["bwd"]["GBps"] / peaks["hbm_gbs"]
    # whole-step roofline of BASELINE.md section 3 (cfg1: 356.1 GF + 543.4 MB + 18.4 MB -> 340 us on 1 GPU)
    t_roof_ms = (356.1e9 / (peak_tc * 1e12) + (543.4e6 + 18.4e6) / (peaks["hbm_gbs"] * 1e9)) * 1e3
    step_frac = t_roof_ms / (ms_value / args.steps)

    # the reference-precision mode (fp32 I/O, hi + lo operand planes: 3 MMAs per k-step) in the same job
    fp32_mode = None
    if rank == 0 and world == 1 and args.precision == "bf16" and not args.no_fp32:
        fp32_mode = measure_fp32_mode(bank, c, n_img, dev, peaks)

    # BASELINE config 5 on the same ranks (forward only, batch sharded, no collective): a short sweep so that the
    # inference row is on the same record as the training line at every N the driver runs
    l2_policy = (f"inputs larger than L2: {n_img}-image feature bank ({bank.numel() * 4 / 1e9:.2f} GB fp32"
                 + (f" + its one-off {bank.numel() * 2 / 1e9:.2f} GB bf16 copy, which the gather reads" if eng.bank_bf16 is not None else "")
                 + ") indexed at random, ~0.6 GB touched per step")
    gru_path = int(lib.vqa_gru_kernel_path())
    gru_label = ({1: "CTA-pair (gru_pair.cu)", 2: "single-CTA (gru.cu)"}.get(gru_path & 3, "?")
                 + (" [pair launch REFUSED earlier]" if gru_path & 256 else ""))
    inference = None
    if not args.no_infer:
        model.engine.close()
        del bank
        torch.cuda.empty_cache()
        sweep, _ = infer_sweep(dp, dev, args.precision, [512, 4096, 8192], 5, 3)
        inference = {"config": "cfg5: forward, K 100 padded boxes (10..100 valid), global batch sharded over the ranks, no collective; "
                               "roofline 0.59 us / sample / GPU", "sweep": sweep}

    # BASELINE config 4 (the vlmap pre-training graph) on the same ranks: a short run of the data-parallel train step so
    # that its row is on the same record as the headline at every N (`--mode memft` prints its own full line)
    pretrain = None
    if not args.no_memft:
        pretrain = measure_memft(dp, dev, args.precision, peaks, steps=5, warmup=3)

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        rate, threads, busy = cpu_port_rate(B, 0, 1, budget_s=10.0)
        cpu_baseline = {"value": rate, "unit": "samples/s", "cores": threads, "kind": "port",
                        "sample": f"fwd+bwd of the oracle's torch-CPU port at cfg1 layer sizes and cfg1's batch ({B}), "
                                  f"{busy:.1f} s of timed CPU work (median step); the reference's TF-1.6 CPU "
                                  f"path cannot run in this image"}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "samples/s", "n_gpus": world, "steps": args.steps,
            "warmup": W, "ms_per_step": ms_value / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16" if args.precision == "bf16" else "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "per_gpu_batch": B, "global_batch": B * world,
                       "precision": args.precision, "parallelism": f"dp{world}", "gradient_collective": collective,
                       "dp_step_check": dp_check,
                       "gru_kernels": gru_label, "l2_policy": l2_policy},
            "e2e": {"value": e2e_value, "unit": "samples/s", "ms_per_step": ms_e2e / args.steps,
                    "h2d_bytes_per_step": h2d * world, "d2h_bytes_per_step": d2h * world,
                    "inputs": ("dense [B, A] soft-score target from pinned host memory" if args.e2e_dense else
                               "ids / tokens / lengths from pinned host memory, soft-score target as sparse (row, id, score) "
                               "triples densified on the device (the native reader's batch format)")},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": roofline,
            "step_roofline": {"t_roof_ms": t_roof_ms, "frac": step_frac,
                              "model": "356.1 GF / bf16 sustained + 561.8 MB / HBM (BASELINE.md section 3)"},
            "attn_hbm": attn,
            "sections": sections,
            "fp32_mode": fp32_mode,
            "inference": inference,
            "pretrain_cfg4": pretrain,
            "phase_ms": phase_ms,
            "critical_path_ms": critical_ms,
            "cpu_baseline": cpu_baseline,
        }
        _emit(line)
    dp.close()
    return 0


def _emit(line):
    """The ONE JSON line goes to the real stdout; everything libraries print (NCCL's version banner, warnings)
    was diverted to stderr at start-up."""
    os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())


_REAL_STDOUT = os.dup(1)
os.dup2(2, 1)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--bank-images", type=int, default=4096)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-fp32", action="store_true", help="skip the fp32-mode measurement (N = 1 only)")
    ap.add_argument("--no-infer", action="store_true", help="skip the BASELINE config 5 inference sweep")
    ap.add_argument("--no-memft", action="store_true", help="skip the short BASELINE config 4 (pre-training graph) run")
    ap.add_argument("--e2e-dense", action="store_true", help="e2e leg uploads the dense [B, A] soft-score target instead of its sparse triples")
    ap.add_argument("--mode", default="train", choices=["train", "infer", "memft"])
    ap.add_argument("--infer-batches", default="64,512,4096,8192")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    if args.mode == "infer":
        return run_infer(args)
    if args.mode == "memft":
        return run_memft(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
