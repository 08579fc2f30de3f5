"""vqa_ops_linear_ln (csrc/linear_ln.cu): one fc_layer -- product, bias, row LayerNorm, activation, Hadamard partner, dropout --
as one cluster kernel, forward and backward, at the shapes the answer model runs it (q_linear_v / q_linear_l / pooled_linear_l:
N 1024 = 16 CTAs x 64 columns; joint_fc: N 2048 = 16 x 128; the data gradient against the answer weights: K 3000, a partial
k-block) plus ragged row counts and the smaller cluster sizes. The checker is torch in fp64 on the same bf16-rounded operands
(vlmap/modules.py:616-650 restated: tf.nn.moments statistics, eps 1e-12); every product call goes through the C ABI."""
import ctypes as C

import numpy as np
import pytest

torch = pytest.importorskip("torch")

from vqa_transfer_externaldata_b200 import lib as L  # noqa: E402

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops():
    lib = L.load()
    h = C.c_void_p()
    L.check(lib.vqa_ops_create(C.byref(h)))
    yield lib, h
    lib.vqa_ops_destroy(h)


def _rand(rng, shape, scale=1.0, shift=0.0):
    return torch.as_tensor(rng.standard_normal(shape) * scale + shift).to("cuda", torch.float32)


def _mask(lib, M, N, keep, seed, step, site):
    m = torch.empty(M * N, dtype=torch.uint8, device="cuda")
    L.check(lib.vqa_ops_dropout_mask(m.data_ptr(), M * N, keep, seed, step, site, None))
    return m.view(M, N).double()


def _tail(z, gamma, beta, mul, mask, keep, act):
    mu = z.mean(dim=1, keepdim=True)
    var = ((z - mu) ** 2).mean(dim=1, keepdim=True)
    rstd = 1.0 / torch.sqrt(var + 1e-12)
    pre = gamma.double() * (z - mu) * rstd + beta.double()
    y = torch.tanh(pre) if act == 1 else torch.relu(pre)
    out = y
    if mul is not None:
        out = out * mul.double()
    if mask is not None:
        out = out * mask / keep
    return mu, rstd, pre, y, out


@pytest.mark.parametrize("M,N,K,act,keep,with_mul", [
    (512, 1024, 2048, 0, 1.0, True),    # pooled_linear_l (x) q_linear_l at cfg1
    (512, 2048, 1024, 0, 0.5, False),   # joint_fc: dropout 0.5
    (48, 1024, 1024, 1, 1.0, False),    # q_L_ft2 (tanh), one ragged row tile
    (200, 512, 256, 0, 0.8, True),      # 8-CTA clusters, ragged last tile
    (130, 64, 64, 0, 1.0, False),       # a cluster of one
    (1000, 1024, 512, 0, 1.0, False),   # more row tiles than co-resident clusters
])
def test_linear_ln_forward(ops, M, N, K, act, keep, with_mul):
    lib, h = ops
    rng = np.random.default_rng(M + N + K)
    a = _rand(rng, (M, K), 0.7).bfloat16()
    w = _rand(rng, (K, N), 1.0 / np.sqrt(K)).bfloat16()
    bias = _rand(rng, (N,), 0.3)
    gamma = _rand(rng, (N,), 0.2, 1.0)
    beta = _rand(rng, (N,), 0.2)
    mul = _rand(rng, (M, N)) if with_mul else None
    z = torch.full((M, N), float("nan"), device="cuda")
    y, out = torch.full_like(z, float("nan")), torch.full_like(z, float("nan"))
    out_hi = torch.zeros(M, N, dtype=torch.bfloat16, device="cuda")
    mean, rstd = torch.zeros(M, device="cuda"), torch.zeros(M, device="cuda")
    seed, step, site = 1234, 7, 3
    d = L.VqaLinearLn(M=M, N=N, K=K, backward=0, a=a.data_ptr(), lda=K, w=w.data_ptr(), ldw=N, bias=bias.data_ptr(),
                      gamma=gamma.data_ptr(), beta=beta.data_ptr(), mul=mul.data_ptr() if with_mul else None, act=act,
                      keep=keep, seed=seed, step=step, site=site, z=z.data_ptr(), mean=mean.data_ptr(), rstd=rstd.data_ptr(),
                      y=y.data_ptr(), out_f32=out.data_ptr(), out_hi=out_hi.data_ptr())
    L.check(lib.vqa_ops_linear_ln(h, C.byref(d), None))
    torch.cuda.synchronize()
    z_ref = a.double() @ w.double() + bias.double()
    mask = _mask(lib, M, N, keep, seed, step, site) if keep < 1.0 else None
    mu_r, rstd_r, pre, y_r, out_r = _tail(z_ref, gamma, beta, mul, mask, keep, act)
    assert torch.allclose(z.double(), z_ref, rtol=0, atol=2e-5 * float(z_ref.abs().max()) + 1e-5)
    assert torch.allclose(mean.double(), mu_r[:, 0], rtol=0, atol=2e-5)
    assert torch.allclose(rstd.double(), rstd_r[:, 0], rtol=2e-5, atol=0)
    near = pre.abs() < 1e-4 if act == 0 else torch.zeros_like(pre, dtype=torch.bool)   # ReLU kinks: either side is right
    assert torch.allclose(y.double()[~near], y_r[~near], rtol=1e-4, atol=1e-4)
    assert torch.allclose(out.double()[~near], out_r[~near], rtol=1e-4, atol=1e-4 * max(1.0, 1.0 / keep))
    assert torch.allclose(out_hi.double()[~near], out_r[~near], rtol=2.0 ** -7, atol=2e-4)
    if mask is not None:   # dropped elements are exact zeros, kept ones are not scaled twice
        assert bool((out[mask == 0] == 0).all())


@pytest.mark.parametrize("M,N,K,act,keep,with_mul,parts", [
    (512, 2048, 3000, 0, 0.5, False, False),   # d logits -> joint_fc backward: the answer dimension's partial k-block
    (512, 1024, 2048, 0, 1.0, True, True),     # d Zj -> pooled_linear_l backward, Hadamard partner, LayerNorm parameter terms
    (77, 1024, 2048, 1, 1.0, False, True),     # tanh, ragged tile
    (300, 256, 200, 0, 0.7, True, False),      # 4-CTA clusters, K below one multi-k-block box
])
def test_linear_ln_backward(ops, M, N, K, act, keep, with_mul, parts):
    lib, h = ops
    rng = np.random.default_rng(3 * M + N + K)
    dy = _rand(rng, (M, K), 0.05).bfloat16()
    w = _rand(rng, (N, K), 1.0 / np.sqrt(N)).bfloat16()
    z = _rand(rng, (M, N), 1.3, 0.2)
    gamma = _rand(rng, (N,), 0.2, 1.0)
    beta = _rand(rng, (N,), 0.2)
    mul = _rand(rng, (M, N)) if with_mul else None
    seed, step, site = 99, 12, 5
    mask = _mask(lib, M, N, keep, seed, step, site) if keep < 1.0 else None
    zr = z.double().requires_grad_(True)
    mu_r, rstd_r, pre, y_r, out_r = _tail(zr, gamma, beta, mul, mask, keep, act)
    raw_ref = dy.double() @ w.double().t()
    out_r.backward(raw_ref)
    mean = mu_r[:, 0].detach().float().contiguous()
    rstd = rstd_r[:, 0].detach().float().contiguous()
    raw = torch.full((M, N), float("nan"), device="cuda")
    dz, dg, db = torch.full_like(raw, float("nan")), torch.full_like(raw, float("nan")), torch.full_like(raw, float("nan"))
    dz_hi = torch.zeros(M, N, dtype=torch.bfloat16, device="cuda")
    d = L.VqaLinearLn(M=M, N=N, K=K, backward=1, a=dy.data_ptr(), lda=K, w=w.data_ptr(), ldw=K, gamma=gamma.data_ptr(),
                      beta=beta.data_ptr(), mul=mul.data_ptr() if with_mul else None, act=act, keep=keep, seed=seed, step=step,
                      site=site, z=z.data_ptr(), mean=mean.data_ptr(), rstd=rstd.data_ptr(), raw=raw.data_ptr(),
                      dz_f32=dz.data_ptr(), dz_hi=dz_hi.data_ptr(), dgamma_part=dg.data_ptr() if parts else None,
                      dbeta_part=db.data_ptr() if parts else None)
    L.check(lib.vqa_ops_linear_ln(h, C.byref(d), None))
    torch.cuda.synchronize()
    scale = float(raw_ref.abs().max())
    assert torch.allclose(raw.double(), raw_ref, rtol=0, atol=2e-5 * scale)
    dz_ref = zr.grad
    near = (pre.detach().abs() < 1e-4).any(dim=1) if act == 0 else torch.zeros(M, dtype=torch.bool, device="cuda")
    tol = 5e-5 * float(dz_ref.abs().max())
    assert torch.allclose(dz.double()[~near], dz_ref[~near], rtol=1e-4, atol=tol)
    assert torch.allclose(dz_hi.double()[~near], dz_ref[~near], rtol=2.0 ** -7, atol=2 * tol)
    if parts:
        dpre = raw_ref * (mul.double() if with_mul else 1.0)
        dpre = dpre * (1.0 - torch.tanh(pre.detach()) ** 2 if act == 1 else (pre.detach() > 0).double())
        xh = (z.double() - mu_r.detach()) * rstd_r.detach()
        nk = pre.detach().abs() >= 1e-4 if act == 0 else torch.ones_like(dpre, dtype=torch.bool)
        assert torch.allclose(db.double()[nk], dpre[nk], rtol=1e-4, atol=2e-5 * scale)
        assert torch.allclose(dg.double()[nk], (dpre * xh)[nk], rtol=1e-4, atol=1e-4 * scale)


def test_linear_ln_refuses_what_it_cannot_run(ops):
    lib, h = ops
    t = torch.zeros(64, 96, device="cuda")
    a = torch.zeros(64, 64, dtype=torch.bfloat16, device="cuda")
    d = L.VqaLinearLn(M=64, N=96, K=64, backward=0, a=a.data_ptr(), lda=64, w=a.data_ptr(), ldw=96, bias=t.data_ptr(),
                      gamma=t.data_ptr(), beta=t.data_ptr(), act=0, keep=1.0, z=t.data_ptr(), mean=t.data_ptr(), rstd=t.data_ptr())
    assert lib.vqa_ops_linear_ln(h, C.byref(d), None) == L.VQA_ERR_BAD_SHAPE   # N = 96: no power-of-two cluster of 64 / 128 columns
    assert b"not eligible" in lib.vqa_last_error()
