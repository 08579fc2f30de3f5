"""Operator-level checks of include/vqa_memft.h at the shapes BASELINE config 4 runs them (the graph-level parity tests in
test_memft_gpu.py use small layers): slab LayerNorm forward / backward on every launch plan (register-resident with one
and with several row groups, shared-memory fallback), the 6-feature weight gradient, softmax cross-entropy gradients and
the spatial attention block with the generic (n = 8) instantiation. The checker is torch autograd in fp64 on the same
device -- test infrastructure, like the oracle; every product call goes through the C ABI."""
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


def _dev(a, dt=torch.float32):
    return torch.as_tensor(np.ascontiguousarray(a)).to("cuda", dt)


def _ref_slab(z, gamma, beta, mul, keep_mask, keep, act):
    z = z.double().requires_grad_(True)
    S, n, N = z.shape
    mu = z.mean(dim=(1, 2), keepdim=True)
    var = ((z - mu) ** 2).mean(dim=(1, 2), keepdim=True)
    pre = gamma.double() * (z - mu) / torch.sqrt(var + 1e-12) + beta.double()
    y = torch.relu(pre) if act == 0 else (torch.tanh(pre) if act == 1 else pre)
    out = y
    if mul is not None:
        out = out * mul.double()
    if keep_mask is not None:
        out = out * keep_mask.double() / keep
    return z, y, out


@pytest.mark.parametrize("slabs,n,N,act,keep,with_mul,smem", [
    (64, 36, 1024, 0, 1.0, False, False),   # spat_v_linear_v at cfg4: four row groups, 48 KB of combine buffer
    (96, 5, 2048, 0, 0.5, False, False),    # joint_fc: dropout
    (96, 5, 1024, 0, 1.0, True, False),     # q_linear_l x v_linear_l
    (40, 5, 1024, 1, 1.0, False, False),    # wordset_ft: tanh
    (32, 36, 1024, 0, 1.0, False, True),    # shared-memory fallback kernels (VQA_SLAB_SMEM)
    (24, 7, 256, 2, 0.8, True, False),      # nine-rows-per-thread plan with one row group, no activation
])
def test_slab_layer_norm_fwd_bwd(ops, monkeypatch, slabs, n, N, act, keep, with_mul, smem):
    lib, h = ops
    if smem:
        monkeypatch.setenv("VQA_SLAB_SMEM", "1")
    rng = np.random.default_rng(slabs + n + N)
    rows = slabs * n
    z = _dev(rng.standard_normal((slabs, n, N)) * 1.5 + 0.3)
    gamma, beta = _dev(1 + 0.2 * rng.standard_normal(N)), _dev(0.2 * rng.standard_normal(N))
    mul_rows = rows // 2 if with_mul else 0
    mul = _dev(rng.standard_normal((mul_rows, N))) if with_mul else None
    dout = _dev(rng.standard_normal((rows, N)))
    dout2 = _dev(rng.standard_normal((rows, N)))
    mean, rstd = torch.zeros(slabs, device="cuda"), torch.zeros(slabs, device="cuda")
    y, out = torch.zeros(rows, N, device="cuda"), torch.zeros(rows, N, device="cuda")
    out_hi = torch.zeros(rows, N, device="cuda", dtype=torch.bfloat16)
    out_lo = torch.zeros_like(out_hi)
    dz, dmul, part = torch.zeros(rows, N, device="cuda"), torch.zeros(rows, N, device="cuda"), torch.zeros(slabs, 3 * N, device="cuda")
    dz_hi = torch.zeros(rows, N, device="cuda", dtype=torch.bfloat16)
    seed, step, site0, rps = 11, 4, 18, rows // 2
    a = L.VqaSlabLn(slabs=slabs, n=n, N=N, act=act, z=z.data_ptr(), gamma=gamma.data_ptr(), beta=beta.data_ptr(),
                    mul=mul.data_ptr() if with_mul else None, mul_rows=mul_rows, keep=keep, seed=seed, step=step, site0=site0,
                    rows_per_site=rps, mean=mean.data_ptr(), rstd=rstd.data_ptr(), y=y.data_ptr(), out_f32=out.data_ptr(),
                    out_hi=out_hi.data_ptr(), out_lo=out_lo.data_ptr(), dout=dout.data_ptr(), dout2=dout2.data_ptr(),
                    dz_f32=dz.data_ptr(), dz_hi=dz_hi.data_ptr(), dmul=dmul.data_ptr(), part=part.data_ptr())
    L.check(lib.vqa_ops_slab_ln_fwd(h, C.byref(a), None))
    L.check(lib.vqa_ops_slab_ln_bwd(h, C.byref(a), None))
    keep_mask = None
    if keep < 1.0:
        m = torch.zeros(rows * N, dtype=torch.uint8, device="cuda")
        for s in range(2):   # two dropout sites of rows_per_site rows each
            L.check(lib.vqa_ops_dropout_mask(C.c_void_p(m.data_ptr() + s * rps * N), rps * N, keep, seed, step, site0 + s, None))
        keep_mask = m.view(slabs, n, N)
    torch.cuda.synchronize()
    mul3 = None if mul is None else mul.repeat(2, 1).view(slabs, n, N)
    zr, y_ref, out_ref = _ref_slab(z, gamma, beta, mul3, keep_mask, keep, act)
    assert (y.view(slabs, n, N).double() - y_ref).abs().max() < 2e-5
    assert (out.view(slabs, n, N).double() - out_ref).abs().max() < 2e-5 * max(1.0, float(out_ref.detach().abs().max()))
    rec = out_hi.double() + out_lo.double()
    assert (rec - out.double()).abs().max() < 1e-4 * max(1.0, float(out.abs().max()))
    up = (dout + dout2).view(slabs, n, N).double()
    grads = torch.autograd.grad((out_ref * up).sum(), zr)[0]
    scale = float(grads.abs().max())
    assert (dz.view(slabs, n, N).double() - grads).abs().max() < 5e-5 * scale
    assert (dz_hi.view(slabs, n, N).double() - grads).abs().max() < 1e-2 * scale
    # d loss / d mul and the per-slab partials of d gamma / d beta / d bias (= column sums of dz)
    drop = 1.0 if keep_mask is None else keep_mask.double() / keep
    if with_mul:
        assert (dmul.view(slabs, n, N).double() - up * drop * y_ref).abs().max() < 5e-5 * max(1.0, float((up * y_ref).abs().max()))
    p3 = part.view(slabs, 3, N).double()
    assert (p3[:, 2] - grads.sum(dim=1)).abs().max() < 1e-4 * max(1.0, float(grads.sum(dim=1).abs().max()))
    zr2 = z.double()
    mu = zr2.mean(dim=(1, 2), keepdim=True)
    xhat = (zr2 - mu) / torch.sqrt(((zr2 - mu) ** 2).mean(dim=(1, 2), keepdim=True) + 1e-12)
    gam = gamma.double().clone().requires_grad_(True)
    bet = beta.double().clone().requires_grad_(True)
    pre = gam * xhat + bet
    yy = torch.relu(pre) if act == 0 else (torch.tanh(pre) if act == 1 else pre)
    oo = yy * (1.0 if mul3 is None else mul3.double()) * drop
    gg, gb = torch.autograd.grad((oo * up).sum(), (gam, bet))
    assert (p3[:, 0].sum(0) - gg).abs().max() < 1e-4 * max(1.0, float(gg.abs().max()))
    assert (p3[:, 1].sum(0) - gb).abs().max() < 1e-4 * max(1.0, float(gb.abs().max()))


@pytest.mark.parametrize("rows,N,boxes", [(18432, 1024, 0), (5120, 1024, 1), (77, 64, 0)])
def test_feature_weight_gradient(ops, rows, N, boxes):
    lib, h = ops
    rng = np.random.default_rng(rows)
    feat = _dev(rng.uniform(size=(rows, 4 if boxes else 6)))
    dzf = _dev(rng.standard_normal((rows, N)))
    hi = dzf.to(torch.bfloat16)
    lo = (dzf - hi.float()).to(torch.bfloat16)
    part, out = torch.zeros(296, 6 * N, device="cuda"), torch.zeros(6, N, device="cuda")
    L.check(lib.vqa_ops_feat_wgrad(h, feat.data_ptr(), 6, boxes, hi.data_ptr(), lo.data_ptr(), rows, N, part.data_ptr(), 296,
                                   out.data_ptr(), None))
    torch.cuda.synchronize()
    f6 = feat.double()
    if boxes:
        f6 = torch.cat([f6, (f6[:, 2] - f6[:, 0]).unsqueeze(1), (f6[:, 3] - f6[:, 1]).unsqueeze(1)], dim=1)
    ref = f6.t() @ (hi.double() + lo.double())
    assert (out.double() - ref).abs().max() < 2e-5 * float(ref.abs().max())


def test_softmax_ce_gradient_matches_autograd(ops):
    lib, h = ops
    B, n, A, heads = 7, 5, 4000, 4
    rng = np.random.default_rng(5)
    R = heads * B * n
    logit = _dev(rng.standard_normal((R, A)) * 2)
    fills = _dev(rng.integers(0, A, size=R), torch.int32)
    num = [_dev(rng.integers(1, n + 1, size=B), torch.int32) for _ in range(2)]
    stats, rep = torch.zeros(R, 4, device="cuda"), torch.zeros(16, device="cuda")
    dl = torch.zeros(R, A, device="cuda")
    dhi = torch.zeros(R, A, device="cuda", dtype=torch.bfloat16)
    a = L.VqaSoftmaxCe(heads=heads, B=B, n=n, A=A, top_k=5, logit=logit.data_ptr(), fills=fills.data_ptr(), loss_scale=1.0,
                       stats=stats.data_ptr(), report=rep.data_ptr(), d_logit=dl.data_ptr(), d_hi=dhi.data_ptr())
    for k in range(heads):
        a.num[k] = num[k % 2].data_ptr()
    L.check(lib.vqa_memft_softmax_ce(h, C.byref(a), None))
    torch.cuda.synchronize()
    lg = logit.double().requires_grad_(True)
    total = 0.0
    for k in range(heads):
        rows = slice(k * B * n, (k + 1) * B * n)
        ce = torch.nn.functional.cross_entropy(lg[rows], fills[rows].long(), reduction="none").view(B, n)
        mask = (torch.arange(n, device="cuda").unsqueeze(0) < num[k % 2].long().unsqueeze(1)).double()
        loss = (ce * mask).sum() / mask.sum()
        assert abs(float(rep[3 * k]) - float(loss)) < 1e-5 * float(loss)
        total = total + loss
    assert abs(float(rep[3 * heads]) - float(total)) < 1e-5 * float(total)
    g = torch.autograd.grad(total, lg)[0]
    assert (dl.double() - g).abs().max() < 1e-6 * max(1.0, float(g.abs().max())) + 1e-9
    assert (dhi.double() - g).abs().max() < 1e-2 * float(g.abs().max())
    # a global count (data-parallel runs) rescales the gradient of that head only
    a.count[1] = 2.0 * float(np.minimum(num[1].cpu().numpy(), n).sum())
    L.check(lib.vqa_memft_softmax_ce(h, C.byref(a), None))
    torch.cuda.synchronize()
    rows = slice(B * n, 2 * B * n)
    assert (dl[rows].double() - 0.5 * g[rows]).abs().max() < 1e-6
    assert (dl[:B * n].double() - g[:B * n]).abs().max() < 1e-6


def test_spatial_attention_generic_entries_against_the_oracle():
    """n = 8 entries per kind (the template instantiation beyond the reference's 5), K not a multiple of the warp count."""
    from test_memft_gpu import run_case
    dims = dict(B=5, K=13, n=8, Dv=64, D=128, L=128, W=24, A=200, T=4, Vq=50, Nws=20)
    err, got, ref, grads, ref_g = run_case(dims, "fp32", seed=7)
    assert all(v < 1e-4 for v in err.values()), err
    for k in ("sv_w", "sq_w", "att_w", "sv_gamma", "sq_beta", "pl_w"):
        scale = max(np.abs(ref_g[k]).max(), 1e-20)
        assert np.abs(grads[k] - ref_g[k]).max() < 1e-4 * scale, k
