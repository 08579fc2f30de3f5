"""Parity of the tcgen05 GEMM (csrc/gemm.cu) through the C ABI (vqa_gemm) against torch matmul.

bf16 mode: operands are bf16, so the reference is an fp32/fp64 matmul of the SAME bf16-rounded values:
only the accumulation order differs (tolerance 2e-5 relative to the max |entry|).
fp32 (split) mode: operands are hi+lo bf16 planes of fp32 values; reference is the fp64 matmul of the
fp32 values; the scheme drops the lo*lo term and the residual beyond 16 mantissa bits: tolerance 5e-5.
"""
import ctypes as C

import pytest

torch = pytest.importorskip("torch")

from vqa_transfer_externaldata_b200 import lib as L  # noqa: E402

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def handle():
    lib = L.load()
    cfg = L.VqaConfig(B=8, K=4, Dv=64, D=64, L=64, J=128, A=64, T=4, W=20, Vq=50, num_train_answer=40,
                      variant=0, precision=0, keep_att=0.8, keep_joint=0.5)
    h = C.c_void_p()
    L.check(lib.vqa_create(C.byref(cfg), C.byref(h)))
    # a workspace gives the CTA-pair kernel its split-K hand-over semaphores
    nbytes = C.c_uint64()
    L.check(lib.vqa_workspace_bytes(h, C.byref(nbytes)))
    ws = torch.empty(nbytes.value + 256, dtype=torch.uint8, device="cuda")
    base = (ws.data_ptr() + 255) // 256 * 256
    L.check(lib.vqa_set_workspace(h, C.c_void_p(base), C.c_uint64(nbytes.value)))
    yield lib, h
    lib.vqa_destroy(h)
    del ws


def _planes(x, split):
    hi = x.to(torch.bfloat16)
    lo = (x - hi.float()).to(torch.bfloat16) if split else None
    return hi, lo


def _run(lib, h, M, N, K, a_mn, b_mn, split, block_n=0, bias=False, addend=False, seed=0, want_bf=True):
    g = torch.Generator(device="cuda").manual_seed(seed)
    A = torch.randn(M, K, device="cuda", generator=g)
    B = torch.randn(N, K, device="cuda", generator=g)
    # pad pitches to multiples of 8 elements (TMA needs 16-byte pitches)
    def store(x, mn_major):
        x = x.t().contiguous() if mn_major else x.contiguous()
        rows, cols = x.shape
        ld = (cols + 7) // 8 * 8
        buf = torch.zeros(rows, ld, device="cuda")
        buf[:, :cols] = x
        return buf, ld
    Ab, lda = store(A, a_mn)
    Bb, ldb = store(B, b_mn)
    a_hi, a_lo = _planes(Ab, split)
    b_hi, b_lo = _planes(Bb, split)
    out = torch.full((M, N), float("nan"), device="cuda")
    out_hi = torch.zeros(M, N, device="cuda", dtype=torch.bfloat16)
    out_lo = torch.zeros(M, N, device="cuda", dtype=torch.bfloat16)
    bias_t = torch.randn(N, device="cuda", generator=g) if bias else None
    add_t = torch.randn(M, N, device="cuda", generator=g) if addend else None
    d = L.VqaGemmDesc(a_hi=a_hi.data_ptr(), a_lo=a_lo.data_ptr() if split else None,
                      b_hi=b_hi.data_ptr(), b_lo=b_lo.data_ptr() if split else None,
                      lda=lda, ldb=ldb, a_mn_major=int(a_mn), b_mn_major=int(b_mn), M=M, N=N, K=K,
                      bias=bias_t.data_ptr() if bias else None,
                      addend=add_t.data_ptr() if addend else None, ld_addend=N,
                      out_f32=out.data_ptr(), ld_f32=N, out_hi=out_hi.data_ptr() if want_bf else None,
                      out_lo=out_lo.data_ptr() if split else None, ld_bf=N, block_n=block_n)
    L.check(lib.vqa_gemm(h, C.byref(d), None))
    torch.cuda.synchronize()
    if split:
        ref = A.double() @ B.double().t()
    else:
        ref = A.to(torch.bfloat16).double() @ B.to(torch.bfloat16).double().t()
    if bias:
        ref = ref + bias_t.double()
    if addend:
        ref = ref + add_t.double()
    scale = ref.abs().max().item()
    err = (out.double() - ref).abs().max().item() / scale
    # bf16 output planes reproduce the fp32 output
    rec = out_hi.double() + (out_lo.double() if split else 0)
    err_planes = (rec - out.double()).abs().max().item() / scale if want_bf else 0.0
    return err, err_planes


LAYOUTS = [(False, False), (False, True), (True, False), (True, True)]


@pytest.mark.parametrize("a_mn,b_mn", LAYOUTS)
@pytest.mark.parametrize("shape", [(128, 128, 64), (256, 256, 256), (512, 1024, 1024), (300, 200, 300),
                                   (130, 3000, 520), (512, 2048, 3000), (200, 320, 456)])   # K = 3000 / 456: partial last k-block
def test_gemm_bf16_layouts(handle, shape, a_mn, b_mn):
    lib, h = handle
    M, N, K = shape
    err, err_p = _run(lib, h, M, N, K, a_mn, b_mn, split=False)
    assert err < 2e-5, (shape, a_mn, b_mn, err)
    assert err_p < 4e-3  # bf16 rounding of the output


@pytest.mark.parametrize("block_n", [64, 128, 256])
@pytest.mark.parametrize("a_mn,b_mn", LAYOUTS)
def test_gemm_bf16_block_n(handle, block_n, a_mn, b_mn):
    lib, h = handle
    err, _ = _run(lib, h, 384, 512, 448, a_mn, b_mn, split=False, block_n=block_n, bias=True, addend=True)
    assert err < 2e-5, (block_n, a_mn, b_mn, err)


@pytest.mark.parametrize("a_mn,b_mn", LAYOUTS)
@pytest.mark.parametrize("shape", [(256, 256, 256), (512, 1024, 2048), (300, 200, 300)])
def test_gemm_split_fp32(handle, shape, a_mn, b_mn):
    lib, h = handle
    M, N, K = shape
    err, err_p = _run(lib, h, M, N, K, a_mn, b_mn, split=True, bias=True)
    assert err < 5e-5, (shape, a_mn, b_mn, err)
    assert err_p < 5e-5


# CTA-pair kernel (csrc/gemm_pair.cu): 256 x |block_n| tiles, persistent; row / column / k tails, all operand
# layouts, bias + addend, bf16 output
@pytest.mark.parametrize("a_mn,b_mn", LAYOUTS)
@pytest.mark.parametrize("shape", [(256, 256, 256), (512, 1024, 1024), (600, 1000, 520), (300, 200, 300),
                                   (130, 3000, 520), (4608, 512, 320)])
@pytest.mark.parametrize("block_n", [-256, -128])
def test_gemm_pair_layouts(handle, shape, a_mn, b_mn, block_n):
    lib, h = handle
    M, N, K = shape
    err, err_p = _run(lib, h, M, N, K, a_mn, b_mn, split=False, block_n=block_n, bias=True, addend=True)
    assert err < 2e-5, (shape, a_mn, b_mn, block_n, err)
    assert err_p < 4e-3


# split-K with the fixed-order hand-over (fp32 output only): few tiles, long K -- the weight-gradient shapes
@pytest.mark.parametrize("a_mn,b_mn", [(True, True), (False, False)])
@pytest.mark.parametrize("shape", [(1024, 512, 4096), (300, 2048, 1792), (512, 3000, 2048), (2048, 1024, 4608)])
def test_gemm_pair_split_k(handle, shape, a_mn, b_mn):
    lib, h = handle
    M, N, K = shape
    err, _ = _run(lib, h, M, N, K, a_mn, b_mn, split=False, block_n=-256, bias=True, addend=True, want_bf=False)
    assert err < 2e-5, (shape, a_mn, b_mn, err)
    # the reduction order is fixed: two runs give the same bits
    outs = []
    for _ in range(2):
        g = torch.Generator(device="cuda").manual_seed(5)
        A = torch.randn(M, K, device="cuda", generator=g).to(torch.bfloat16)
        B = torch.randn(N, K, device="cuda", generator=g).to(torch.bfloat16)
        out = torch.empty(M, N, device="cuda")
        d = L.VqaGemmDesc(a_hi=A.data_ptr(), b_hi=B.data_ptr(), lda=K, ldb=K, M=M, N=N, K=K,
                          out_f32=out.data_ptr(), ld_f32=N, block_n=-256)
        L.check(lib.vqa_gemm(h, C.byref(d), None))
        torch.cuda.synchronize()
        outs.append(out.clone())
    assert torch.equal(outs[0], outs[1])


def test_gemm_rejects_bad_pitch(handle):
    lib, h = handle
    a = torch.zeros(64, 64, device="cuda", dtype=torch.bfloat16)
    o = torch.zeros(64, 64, device="cuda")
    d = L.VqaGemmDesc(a_hi=a.data_ptr(), b_hi=a.data_ptr(), lda=63, ldb=64, M=64, N=64, K=64,
                      out_f32=o.data_ptr(), ld_f32=64)
    assert lib.vqa_gemm(h, C.byref(d), None) == L.VQA_ERR_BAD_SHAPE
    assert b"lda" in lib.vqa_last_error()
