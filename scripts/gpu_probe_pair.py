"""Separate the fixed cost, the per-k-block cost and the epilogue cost of the CTA-pair GEMM (run on the GPU box)."""
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
nbytes = C.c_uint64()
L.check(lib.vqa_workspace_bytes(h, C.byref(nbytes)))
ws = torch.empty(nbytes.value + 256, dtype=torch.uint8, device="cuda")
L.check(lib.vqa_set_workspace(h, C.c_void_p((ws.data_ptr() + 255) // 256 * 256), C.c_uint64(nbytes.value)))
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def bench(name, M, N, K, a_mn, b_mn, block_n, out="f32", iters=15, do_flush=True):
    A = torch.randn((K, M) if a_mn else (M, K), device="cuda").to(torch.bfloat16)
    B = torch.randn((K, N) if b_mn else (N, K), device="cuda").to(torch.bfloat16)
    o32 = torch.empty(M, N, device="cuda") if out == "f32" else None
    o16 = torch.empty(M, N, device="cuda", dtype=torch.bfloat16) if out == "bf16" else None
    d = L.VqaGemmDesc(a_hi=A.data_ptr(), b_hi=B.data_ptr(), lda=A.shape[1], ldb=B.shape[1],
                      a_mn_major=int(a_mn), b_mn_major=int(b_mn), M=M, N=N, K=K,
                      out_f32=o32.data_ptr() if o32 is not None else None, ld_f32=N,
                      out_hi=o16.data_ptr() if o16 is not None else None, ld_bf=N, block_n=block_n)
    for _ in range(3):
        L.check(lib.vqa_gemm(h, C.byref(d), None))
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        if do_flush:
            flush.zero_()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        L.check(lib.vqa_gemm(h, C.byref(d), None))
        e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    t = ts[len(ts) // 2]
    fl = 2.0 * M * N * K
    print(f"{name:26s} M={M:6d} N={N:5d} K={K:6d} mn={int(a_mn)}{int(b_mn)} bn={block_n:4d} out={out:4s} flush={int(do_flush)} "
          f"{t*1e3:8.1f} us  {fl/t/1e9:8.1f} TFLOP/s", flush=True)


P = 74
if os.environ.get("HEADS"):
    for (M, N, K, amn, bmn) in ((512, 1024, 1024, False, True), (512, 1024, 2048, False, True), (512, 2048, 1024, False, True),
                                (512, 3008, 2048, False, True), (512, 2048, 3008, False, False), (512, 1024, 2048, False, False)):
        bench("head auto", M, N, K, amn, bmn, 0, do_flush=False)
        bench("head 128x64 per-kb", M, N, K, amn, bmn, 64, do_flush=False)
    sys.exit(0)
if os.environ.get("QUICK"):
    bench("one round", 256 * P, 256, 64, False, True, -256)
    bench("four rounds", 256 * P, 1024, 64, False, True, -256)
    bench("one round bf16", 256 * P, 256, 64, False, True, -256, out="bf16")
    bench("tiny", 256, 256, 64, False, True, -256)
    sys.exit(0)
for K in (64, 320, 1024, 2048, 4096):
    bench("one round", 256 * P, 256, K, False, True, -256)
for K in (64, 2048):
    bench("one round bf16", 256 * P, 256, K, False, True, -256, out="bf16")
    bench("two rounds", 256 * P, 512, K, False, True, -256)
    bench("four rounds", 256 * P, 1024, K, False, True, -256)
bench("one round warm", 256 * P, 256, 64, False, True, -256, do_flush=False)
bench("xg", 7168, 2048, 320, False, True, -256)
bench("xg old", 7168, 2048, 320, False, True, 256)
bench("xg warm", 7168, 2048, 320, False, True, -256, do_flush=False)
bench("xg bf16", 7168, 2048, 320, False, True, -256, out="bf16")
bench("logits split", 512, 3000, 2048, False, True, -256)
bench("logits old", 512, 3000, 2048, False, True, 64)
bench("dWg_h split", 1024, 2048, 7168, True, True, -256)
bench("dWg_h old", 1024, 2048, 7168, True, True, 64)
bench("vproj wgrad", 2048, 1024, 18432, True, True, -256)
bench("vproj fwd", 18432, 1024, 2048, False, True, -256, out="bf16")
