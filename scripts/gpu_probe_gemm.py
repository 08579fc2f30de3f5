"""Time the tcgen05 GEMM on the shapes of the answer model (run on the GPU box)."""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vqa_transfer_externaldata_b200 import lib as L

lib = L.load()
cfg = L.VqaConfig(B=8, K=4, Dv=64, D=64, L=64, J=128, A=64, T=4, W=20, Vq=50, num_train_answer=40,
                  variant=0, precision=0, keep_att=0.8, keep_joint=0.5)
h = C.c_void_p()
L.check(lib.vqa_create(C.byref(cfg), C.byref(h)))
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def bench(name, M, N, K, a_mn, b_mn, split, block_n=0, iters=20):
    A = torch.randn((K, M) if a_mn else (M, K), device="cuda")
    B = torch.randn((K, N) if b_mn else (N, K), device="cuda")
    a_hi = A.to(torch.bfloat16); a_lo = (A - a_hi.float()).to(torch.bfloat16)
    b_hi = B.to(torch.bfloat16); b_lo = (B - b_hi.float()).to(torch.bfloat16)
    out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    d = L.VqaGemmDesc(a_hi=a_hi.data_ptr(), a_lo=a_lo.data_ptr() if split else None,
                      b_hi=b_hi.data_ptr(), b_lo=b_lo.data_ptr() if split else None,
                      lda=A.shape[1], ldb=B.shape[1], a_mn_major=int(a_mn), b_mn_major=int(b_mn),
                      M=M, N=N, K=K, out_hi=out.data_ptr(), ld_bf=N, block_n=block_n)
    for _ in range(3):
        L.check(lib.vqa_gemm(h, C.byref(d), None))
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        flush.zero_()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        L.check(lib.vqa_gemm(h, C.byref(d), None))
        e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    t = ts[len(ts) // 2]
    fl = 2.0 * M * N * K * (3 if split else 1)
    print(f"{name:28s} M={M:6d} N={N:5d} K={K:6d} a_mn={int(a_mn)} b_mn={int(b_mn)} split={int(split)} "
          f"bn={block_n:3d}  {t*1e3:8.1f} us  {fl/t/1e9:8.1f} TFLOP/s(mma)", flush=True)


for bn in (128, 256):
    bench("vproj fwd", 18432, 1024, 2048, False, True, False, bn)
    bench("vproj wgrad", 2048, 1024, 18432, True, True, False, bn)
bench("vproj fwd fp32", 18432, 1024, 2048, False, True, True, 128)
bench("vproj wgrad fp32", 2048, 1024, 18432, True, True, True, 128)
for bn in (64, 128):
    bench("gru gates step", 512, 2048, 1024, False, True, False, bn)
    bench("gru cand step", 512, 1024, 1024, False, True, False, bn)
    bench("answer head", 512, 3000, 2048, False, True, False, bn)
    bench("answer dgrad", 512, 2048, 3000, False, False, False, bn)
bench("gru wgrad", 1024, 2048, 7168, True, True, False, 128)
bench("square 8192", 8192, 8192, 8192, False, False, False, 256)
bench("square 8192", 8192, 8192, 8192, False, False, False, 128)
