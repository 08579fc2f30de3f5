"""Where do ~20 us go in a 2 GFLOP GEMM? Back-to-back launches, swept over K, tile width and epilogue."""
import ctypes as C, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vqa_transfer_externaldata_b200 import lib as L
lib = L.load()
cfg = L.VqaConfig(B=8, K=4, Dv=64, D=64, L=64, J=128, A=64, T=4, W=20, Vq=50, num_train_answer=40,
                  variant=0, precision=0, keep_att=0.8, keep_joint=0.5)
h = C.c_void_p(); L.check(lib.vqa_create(C.byref(cfg), C.byref(h)))

def bench(M, N, K, bn, b_mn, addend, out, reps=200):
    A = torch.randn(M, K, device="cuda").to(torch.bfloat16)
    B = torch.randn((K, N) if b_mn else (N, K), device="cuda").to(torch.bfloat16)
    of = torch.empty(M, N, device="cuda"); ob = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    add = torch.randn(M, N, device="cuda")
    d = L.VqaGemmDesc(a_hi=A.data_ptr(), b_hi=B.data_ptr(), lda=K, ldb=B.shape[1], a_mn_major=0, b_mn_major=int(b_mn),
                      M=M, N=N, K=K, addend=add.data_ptr() if addend else None, ld_addend=N,
                      out_f32=of.data_ptr() if out == "f32" else None, ld_f32=N,
                      out_hi=ob.data_ptr() if out == "bf16" else None, ld_bf=N, block_n=bn)
    for _ in range(5): L.check(lib.vqa_gemm(h, C.byref(d), None))
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): lib.vqa_gemm(h, C.byref(d), None)
    e1.record(); torch.cuda.synchronize()
    t = e0.elapsed_time(e1) / reps * 1e3
    print(f"M={M} N={N} K={K:5d} bn={bn:3d} b_mn={int(b_mn)} addend={int(addend)} out={out:4s}: {t:7.2f} us/launch  "
          f"{2.0*M*N*K/t/1e6:7.1f} TF/s", flush=True)

for K in (64, 256, 1024, 4096):
    bench(512, 2048, K, 64, True, False, "f32")
for bn in (64, 128, 256):
    bench(512, 2048, 1024, bn, True, True, "f32")
    bench(512, 2048, 1024, bn, False, False, "bf16")
bench(512, 1024, 1024, 64, True, True, "f32")
bench(128, 64, 1024, 64, True, False, "f32")     # a single CTA
bench(128, 64, 64, 64, True, False, "f32")       # a single CTA, a single k-block
bench(18432, 1024, 2048, 256, True, False, "bf16", reps=20)
bench(18432, 1024, 2048, 128, True, False, "bf16", reps=20)
