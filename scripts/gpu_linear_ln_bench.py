"""Tile-plan sweep of the fused fc_layer kernel (csrc/linear_ln.cu) on the shapes of the answer model's head, against the
unfused pair (vqa_ops_gemm + row LayerNorm). CUDA events, two regimes per plan: back-to-back launches (operands hot in
L2, prologues overlapped by programmatic dependent launch) and isolated launches after an L2 flush (median).
Usage (GPU box): VQA_LINEAR_LN_TUNE=1 python scripts/gpu_linear_ln_bench.py > gpurun_out/linear_ln_sweep.txt"""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vqa_transfer_externaldata_b200 import lib as L  # noqa: E402

os.environ["VQA_LINEAR_LN_TUNE"] = "1"
lib = L.load()
ops = C.c_void_p()
L.check(lib.vqa_ops_create(C.byref(ops)))
dev = "cuda"
flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timed(fn, reps=30):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    hot = e0.elapsed_time(e1) / reps * 1e3
    iso = []
    for _ in range(9):
        flush_buf.zero_()
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        iso.append(e0.elapsed_time(e1) * 1e3)
    return hot, float(np.median(iso))


def case(M, N, K, backward):
    g = torch.Generator(device=dev).manual_seed(M + N + K)
    a = (torch.randn(M, K, device=dev, generator=g) * 0.3).bfloat16()
    w = (torch.randn((N, K) if backward else (K, N), device=dev, generator=g) * 0.03).bfloat16()
    f = lambda *s: torch.randn(*s, device=dev, generator=g)
    bias, gamma, beta = f(N), f(N) * 0.1 + 1, f(N) * 0.1
    z, y, raw, dz = f(M, N), f(M, N), f(M, N), f(M, N)
    mean, rstd = f(M), f(M).abs() + 0.5
    hi = torch.zeros(M, N, dtype=torch.bfloat16, device=dev)
    d = L.VqaLinearLn(M=M, N=N, K=K, backward=backward, a=a.data_ptr(), lda=K, w=w.data_ptr(), ldw=K if backward else N,
                      bias=bias.data_ptr(), gamma=gamma.data_ptr(), beta=beta.data_ptr(), act=0, keep=0.5, seed=1, step=2, site=3,
                      z=z.data_ptr(), mean=mean.data_ptr(), rstd=rstd.data_ptr(),
                      y=None if backward else y.data_ptr(), out_hi=None if backward else hi.data_ptr(),
                      raw=raw.data_ptr() if backward else None, dz_hi=hi.data_ptr() if backward else None)
    gd = L.VqaGemmDesc(M=M, N=N, K=K, a_hi=a.data_ptr(), lda=K, a_mn_major=0, b_hi=w.data_ptr(), ldb=K if backward else N,
                       b_mn_major=0 if backward else 1, bias=None if backward else bias.data_ptr(), out_f32=raw.data_ptr(), ld_f32=N)
    sl = L.VqaSlabLn(slabs=M, n=1, N=N, act=0, z=raw.data_ptr(), gamma=gamma.data_ptr(), beta=beta.data_ptr(), keep=0.5, seed=1, step=2,
                     site0=3, rows_per_site=M, mean=mean.data_ptr(), rstd=rstd.data_ptr(), y=None if backward else y.data_ptr(),
                     out_hi=hi.data_ptr() if not backward else None, dout=raw.data_ptr() if backward else None,
                     dz_hi=hi.data_ptr() if backward else None)

    def unfused():
        L.check(lib.vqa_ops_gemm(ops, C.byref(gd), 0, None))
        if backward:
            sl.z = z.data_ptr()
            L.check(lib.vqa_ops_slab_ln_bwd(ops, C.byref(sl), None))
        else:
            L.check(lib.vqa_ops_slab_ln_fwd(ops, C.byref(sl), None))

    def gemm_only():
        L.check(lib.vqa_ops_gemm(ops, C.byref(gd), 0, None))

    tag = f"{'bwd' if backward else 'fwd'} M{M} N{N} K{K}"
    print(f"{tag:24s} unfused gemm+ln      hot {timed(unfused)[0]:7.2f} us   isolated {timed(unfused)[1]:7.2f} us", flush=True)
    print(f"{tag:24s} gemm alone           hot {timed(gemm_only)[0]:7.2f} us   isolated {timed(gemm_only)[1]:7.2f} us", flush=True)
    for bn in (64, 128, 256):
        for rows in (64, 128):
            for mc in (1, 0):
                os.environ["VQA_LINEAR_LN_BN"], os.environ["VQA_LINEAR_LN_ROWS"], os.environ["VQA_LINEAR_LN_MC"] = str(bn), str(rows), str(mc)
                if lib.vqa_ops_linear_ln(ops, C.byref(d), None) != 0:
                    continue
                torch.cuda.synchronize()
                hot, iso = timed(lambda: L.check(lib.vqa_ops_linear_ln(ops, C.byref(d), None)))
                print(f"{tag:24s} bn {bn:3d} cl {N // bn:2d} rows {rows:3d} mc {mc}  hot {hot:7.2f} us   isolated {iso:7.2f} us", flush=True)


for shp in ((512, 1024, 1024, 0), (512, 1024, 2048, 0), (512, 2048, 1024, 0), (512, 2048, 3000, 1), (512, 1024, 2048, 1)):
    case(*shp)
