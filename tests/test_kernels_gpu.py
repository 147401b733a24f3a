"""GPU: every C-ABI kernel against a plain torch fp64/fp32 restatement of the same op on the same
seeded inputs.  Tolerances (relative L2): fp32 1e-4 (north_star), bf16 2e-2."""
import pytest
import torch
import torch.nn.functional as F

from conftest import rel_l2

pytestmark = pytest.mark.gpu
TOL = {torch.float32: 1e-4, torch.bfloat16: 2e-2, torch.float16: 5e-3}
DTYPES = [torch.float32, torch.bfloat16, torch.float16]


def backends(vsw):
    return [vsw._lib.GEMM_SIMT, vsw._lib.GEMM_AUTO]


@pytest.fixture(autouse=True)
def _reset_backend(vsw):
    yield
    vsw._lib.set_gemm_backend(vsw._lib.GEMM_AUTO)


def rnd(*shape, dtype=torch.float32, scale=1.0, seed=None):
    if seed is not None:
        torch.manual_seed(seed)
    return (torch.randn(*shape, device="cuda") * scale).to(dtype)


# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("dtype", DTYPES + [torch.float16])
@pytest.mark.parametrize("C", [32, 96, 128, 1024, 3072])
def test_layernorm_fwd_bwd(vsw, dtype, C):
    VF = vsw.functional
    torch.manual_seed(C)
    B, T = 3, 37
    x = rnd(B, T, C, dtype=dtype, scale=2.0) + 0.5
    g = rnd(C, dtype=dtype) * 0.2 + 1
    b = rnd(C, dtype=dtype) * 0.1
    dy = rnd(B, T, C, dtype=dtype)
    dres = rnd(B, T, C, dtype=dtype)
    y, mean, rstd = VF.ln_fwd(x, g, b, None, B, T, T, C)
    xr = x.double().requires_grad_(True)
    gr, br = g.double().requires_grad_(True), b.double().requires_grad_(True)
    yr = F.layer_norm(xr, (C,), gr, br, 1e-5)
    assert rel_l2(y, yr) < TOL[dtype]
    assert rel_l2(mean, x.double().mean(-1).reshape(-1)) < 1e-5
    yr.backward(dy.double())
    dx, dg, db = VF.ln_bwd(dy, x, g, mean, rstd, None, dres, B, T, T, C)
    assert rel_l2(dx, xr.grad + dres.double()) < TOL[dtype]
    assert rel_l2(dg, gr.grad) < 1e-4 and rel_l2(db, br.grad) < 1e-4


@pytest.mark.parametrize("dtype", DTYPES)
def test_layernorm_gather_with_padding(vsw, oracle, dtype):
    """norm1 + zero-pad-after-norm + roll + window_partition as ONE gather-LN kernel, and its transpose"""
    VF = vsw.functional
    grid, window, shift = (4, 11, 13), (8, 7, 7), (4, 3, 3)
    B, C = 2, 64
    T = grid[0] * grid[1] * grid[2]
    plan = VF.window_plan(grid, window, shift, "cuda")
    R = plan.nW * plan.N
    x = rnd(B, T, C, dtype=dtype, seed=1)
    g, b = rnd(C, dtype=dtype) * 0.2 + 1, rnd(C, dtype=dtype) * 0.1
    y, mean, rstd = VF.ln_fwd(x, g, b, plan.gather, B, T, R, C)
    n = F.layer_norm(x.double(), (C,), g.double(), b.double(), 1e-5)
    n = torch.cat([n, n.new_zeros(B, 1, C)], 1)
    ref = n[:, plan.gather.long()]  # -1 -> the appended zero row
    assert rel_l2(y, ref) < TOL[dtype]
    assert float(y[:, plan.gather < 0].abs().max()) == 0.0
    dy = rnd(B, R, C, dtype=dtype)
    dres = rnd(B, T, C, dtype=dtype)
    xr = x.double().requires_grad_(True)
    gr, br = g.double().requires_grad_(True), b.double().requires_grad_(True)
    n = F.layer_norm(xr, (C,), gr, br, 1e-5)
    n = torch.cat([n, n.new_zeros(B, 1, C)], 1)[:, plan.gather.long()]
    n.backward(dy.double())
    dx, dg, db = VF.ln_bwd(dy, x, g, mean, rstd, plan.gather, dres, B, T, R, C)
    assert rel_l2(dx, xr.grad + dres.double()) < TOL[dtype]
    assert rel_l2(dg, gr.grad) < 1e-4 and rel_l2(db, br.grad) < 1e-4


# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("M,N,K", [(392, 96, 32), (1000, 384, 128), (777, 128, 512), (300, 2048, 512), (64, 24, 96)])
@pytest.mark.parametrize("backend", [1, 2])
def test_linear_fwd_epilogues(vsw, dtype, M, N, K, backend):
    VF, L = vsw.functional, vsw._lib
    L.set_gemm_backend(backend)
    x, w, b = rnd(M, K, dtype=dtype, seed=M + N), rnd(N, K, dtype=dtype, scale=0.05), rnd(N, dtype=dtype, scale=0.1)
    ref = x.double() @ w.double().t() + b.double()
    y = VF.linear_fwd(x, w, b, M, N, K)
    assert rel_l2(y, ref) < TOL[dtype]
    y = VF.linear_fwd(x, w, None, M, N, K)
    assert rel_l2(y, x.double() @ w.double().t()) < TOL[dtype]
    u = torch.empty(M, N, dtype=dtype, device="cuda")
    y = VF.linear_fwd(x, w, b, M, N, K, epi=L.EPI_GELU, aux_out=u)
    assert rel_l2(u, ref) < TOL[dtype] and rel_l2(y, F.gelu(ref)) < TOL[dtype]
    # training form: the second output is gelu'(pre-activation)
    gd = torch.empty_like(u)
    y2 = VF.linear_fwd(x, w, b, M, N, K, epi=L.EPI_GELU_GRAD, aux_out=gd)
    rd = ref.clone().requires_grad_(True)
    F.gelu(rd).sum().backward()
    assert rel_l2(y2, F.gelu(ref)) < TOL[dtype] and rel_l2(gd, rd.grad) < TOL[dtype]
    res = rnd(M, N, dtype=dtype)
    y = VF.linear_fwd(x, w, b, M, N, K, epi=L.EPI_RESIDUAL, res=res, rows_per_batch=M, dst_rows_per_batch=M)
    assert rel_l2(y, res.double() + ref) < TOL[dtype]


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("backend", [1, 2])
def test_linear_scatter_residual_epilogue_and_its_transpose(vsw, dtype, backend):
    """proj + window_reverse + roll-back + crop + drop-path + residual in one epilogue; the dgrad
    gathers the same rows"""
    VF, L = vsw.functional, vsw._lib
    L.set_gemm_backend(backend)
    grid, window, shift = (4, 11, 13), (8, 7, 7), (4, 3, 3)
    plan = VF.window_plan(grid, window, shift, "cuda")
    B, C, T, R = 2, 64, 4 * 11 * 13, plan.nW * plan.N
    o = rnd(B * R, C, dtype=dtype, seed=5)
    w, b = rnd(C, C, dtype=dtype, scale=0.1), rnd(C, dtype=dtype, scale=0.1)
    res = rnd(B, T, C, dtype=dtype)
    scale = torch.tensor([0.0, 1.25], device="cuda")
    y = VF.linear_fwd(o, w, b, B * R, C, C, epi=L.EPI_RESIDUAL, res=res, rowmap=plan.gather, rowscale=scale,
                      rows_per_batch=R, dst_rows_per_batch=T, out_rows=B * T).view(B, T, C)
    lin = (o.double() @ w.double().t() + b.double()).view(B, R, C)
    ref = res.double().clone()
    gm = plan.gather.long()
    ok = gm >= 0
    ref[:, gm[ok]] += lin[:, ok] * scale.double().view(B, 1, 1)
    assert rel_l2(y, ref) < TOL[dtype]
    dy = rnd(B, T, C, dtype=dtype)
    a_out = torch.empty(B * R, C, dtype=dtype, device="cuda")
    dx = VF.linear_dgrad(dy, w, B * R, C, C, a_rowmap=plan.gather, a_rowscale=scale, rows_per_batch=R,
                         src_rows_per_batch=T, a_out=a_out)
    a_ref = torch.zeros(B, R, C, dtype=torch.float64, device="cuda")
    a_ref[:, ok] = dy.double()[:, gm[ok]] * scale.double().view(B, 1, 1)
    assert rel_l2(a_out, a_ref.view(B * R, C)) < TOL[dtype]
    assert rel_l2(dx, a_ref.view(B * R, C) @ w.double()) < TOL[dtype]


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("M,N,K", [(392, 96, 32), (3000, 384, 128), (1111, 128, 512), (4096, 512, 2048), (50, 24, 96)])
@pytest.mark.parametrize("backend", [1, 2])
def test_linear_dgrad_wgrad(vsw, dtype, M, N, K, backend):
    VF, L = vsw.functional, vsw._lib
    L.set_gemm_backend(backend)
    dy, w, x = rnd(M, N, dtype=dtype, seed=K), rnd(N, K, dtype=dtype, scale=0.05), rnd(M, K, dtype=dtype)
    u = rnd(M, K, dtype=dtype)
    dx = VF.linear_dgrad(dy, w, M, N, K)
    assert rel_l2(dx, dy.double() @ w.double()) < TOL[dtype]
    dxg = VF.linear_dgrad(dy, w, M, N, K, gelu_pre=u)
    ud = u.double().requires_grad_(True)
    F.gelu(ud).backward(dy.double() @ w.double())
    assert rel_l2(dxg, ud.grad) < TOL[dtype]
    dxm = VF.linear_dgrad(dy, w, M, N, K, mul=u)
    assert rel_l2(dxm, (dy.double() @ w.double()) * u.double()) < TOL[dtype]
    dw, db = VF.linear_wgrad(dy, x, M, N, K)
    assert rel_l2(dw, dy.double().t() @ x.double()) < TOL[dtype]
    assert rel_l2(db, dy.double().sum(0)) < TOL[dtype]
    dw32, _ = VF.linear_wgrad(dy, x, M, N, K, need_bias=False, grad_dtype=torch.float32)
    assert dw32.dtype == torch.float32 and rel_l2(dw32, dy.double().t() @ x.double()) < 1e-4


# ---------------------------------------------------------------------------------------------
def attn_reference(qkv, table, rel_index, mask, nW, nH, scale):
    """plain torch restatement of video_swin.py:149-169 in fp64"""
    B_, N, _, _, hd = qkv.shape
    q, k, v = [qkv[:, :, i].transpose(1, 2).double() for i in range(3)]
    s = (q * scale) @ k.transpose(-1, -2)
    bias = table.double()[rel_index[:N, :N].reshape(-1)].view(N, N, nH).permute(2, 0, 1)
    s = s + bias[None]
    if mask is not None:
        s = (s.view(B_ // nW, nW, nH, N, N) + mask.double()[None, :, None]).view(B_, nH, N, N)
    p = torch.softmax(s, -1)
    return (p @ v).transpose(1, 2).reshape(B_, N, nH * hd), torch.logsumexp(s, -1)


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("geom", [((8, 14, 14), (8, 7, 7), (0, 3, 3), 3), ((8, 7, 7), (8, 7, 7), (0, 0, 0), 2),
                                  ((4, 12, 12), (4, 6, 6), (0, 3, 3), 4), ((16, 7, 7), (8, 7, 7), (4, 0, 0), 1),
                                  ((8, 4, 4), (8, 7, 7), (4, 3, 3), 2)])
@pytest.mark.parametrize("backend", [1, 2])
def test_window_attention_fwd_bwd(vsw, oracle, dtype, geom, backend):
    VF, L = vsw.functional, vsw._lib
    L.set_gemm_backend(backend)
    grid, window, shift, nH = geom
    hd, B = 32, 2
    plan = VF.window_plan(grid, window, shift, "cuda")
    nW, N = plan.nW, plan.N
    B_ = B * nW
    torch.manual_seed(nH)
    qkv = rnd(B_, N, 3, nH, hd, dtype=dtype)
    Lt = (2 * window[0] - 1) * (2 * window[1] - 1) * (2 * window[2] - 1)
    table = rnd(Lt, nH, dtype=dtype, scale=0.5)
    rel_index = torch.from_numpy(oracle.relative_position_index(window)).cuda()
    rowcode, colcode = VF.bias_codes(rel_index, N)
    mask = vsw.compute_mask(*plan.pgrid, plan.ws, plan.ss, "cuda") if plan.shifted else None
    qr = qkv.double().requires_grad_(True)
    tr = table.double().requires_grad_(True)
    oref, lref = attn_reference(qr, tr, rel_index, mask, nW, nH, hd ** -0.5)
    out, lse = VF.attn_fwd(qkv.view(B_ * N, -1), table, rowcode, colcode, plan.region, None, B_, nW, N, nH, hd, hd ** -0.5)
    assert rel_l2(out.view(B_, N, -1), oref) < TOL[dtype]
    assert rel_l2(lse, lref) < (1e-5 if dtype == torch.float32 else 1e-2)
    # with the configured window as layout hint (padded on-chip bias table in the tcgen05 kernels): same result
    out_h, lse_h = VF.attn_fwd(qkv.view(B_ * N, -1), table, rowcode, colcode, plan.region, None, B_, nW, N, nH, hd, hd ** -0.5,
                               window=window)
    assert rel_l2(out_h.view(B_, N, -1), oref) < TOL[dtype] and rel_l2(lse_h, lref) < (1e-5 if dtype == torch.float32 else 1e-2)
    if plan.shifted:  # the dense-mask path must agree with the region-id path
        L.set_gemm_backend(L.GEMM_SIMT)
        out2, _ = VF.attn_fwd(qkv.view(B_ * N, -1), table, rowcode, colcode, None, mask.to(dtype), B_, nW, N, nH, hd, hd ** -0.5)
        assert rel_l2(out2.view(B_, N, -1), oref) < TOL[dtype]
        L.set_gemm_backend(backend)
    dout = rnd(B_, N, nH * hd, dtype=dtype)
    oref.backward(dout.double())
    dqkv, dtab = VF.attn_bwd(qkv.view(B_ * N, -1), out, dout.view(B_ * N, -1), lse, table, rowcode, colcode, plan.region,
                             None, B_, nW, N, nH, hd, hd ** -0.5)
    dq = dqkv.view(B_, N, 3, nH, hd)
    for i, nm in enumerate("qkv"):
        assert rel_l2(dq[:, :, i], qr.grad[:, :, i]) < TOL[dtype] * (1 if dtype == torch.float32 else 1.5), nm
    assert rel_l2(dtab, tr.grad) < TOL[dtype] * (1 if dtype == torch.float32 else 1.5)
    dqkv_h, dtab_h = VF.attn_bwd(qkv.view(B_ * N, -1), out, dout.view(B_ * N, -1), lse, table, rowcode, colcode, plan.region,
                                 None, B_, nW, N, nH, hd, hd ** -0.5, planes=plan.ws[0], window=window)
    # (the hinted call may run a different kernel generation than the un-hinted one: both are held against the fp64 reference)
    dqh = dqkv_h.view(B_, N, 3, nH, hd)
    for i, nm in enumerate("qkv"):
        assert rel_l2(dqh[:, :, i], qr.grad[:, :, i]) < TOL[dtype] * (1 if dtype == torch.float32 else 1.5), nm
    assert rel_l2(dqkv_h, dqkv) < (1e-5 if dtype == torch.float32 else 2 * TOL[dtype])
    assert rel_l2(dtab_h, tr.grad) < TOL[dtype] * (1 if dtype == torch.float32 else 1.5)


@pytest.mark.parametrize("dtype,hint", [(torch.bfloat16, False), (torch.bfloat16, True), (torch.float16, True)])
@pytest.mark.parametrize("case", ["huge_logits", "huge_bias"])
def test_window_attention_exact_softmax_path(vsw, oracle, case, dtype, hint):
    """The tcgen05 forward skips the row-max subtraction only when |scores| and |bias| are provably small (bf16: |exponent| <= 50;
    fp16: Cauchy-Schwarz bound <= 8, then a uniform shift keeps P <= 256); huge logits or bias values must take the exact
    two-pass path and still match (and the recompute backward must accept its log-sum-exp).  hint = the configured window as
    layout hint, i.e. the second-generation kernels (fp16 exists only there); without it the first-generation bf16 kernels."""
    VF, L = vsw.functional, vsw._lib
    wkw = dict(window=(8, 7, 7)) if hint else {}
    L.set_gemm_backend(L.GEMM_TCGEN05)
    try:
        grid, window, shift, nH, hd, B = (8, 14, 14), (8, 7, 7), (0, 3, 3), 2, 32, 1
        plan = VF.window_plan(grid, window, shift, "cuda")
        nW, N = plan.nW, plan.N
        B_ = B * nW
        torch.manual_seed(11)
        qkv = rnd(B_, N, 3, nH, hd, dtype=dtype)
        table = rnd(15 * 13 * 13, nH, dtype=dtype, scale=0.5)
        if case == "huge_logits":
            qkv[:, :, :2] *= 6.0        # |q||k| scale log2e ~ 400 >> 50
        else:
            table = table * 150.0       # bias range far beyond +-50 log2 units
        rel_index = torch.from_numpy(oracle.relative_position_index(window)).cuda()
        rowcode, colcode = VF.bias_codes(rel_index, N)
        mask = vsw.compute_mask(*plan.pgrid, plan.ws, plan.ss, "cuda")
        qr = qkv.double().requires_grad_(True)
        tr = table.double().requires_grad_(True)
        oref, lref = attn_reference(qr, tr, rel_index, mask, nW, nH, hd ** -0.5)
        out, lse = VF.attn_fwd(qkv.view(B_ * N, -1), table, rowcode, colcode, plan.region, None, B_, nW, N, nH, hd, hd ** -0.5,
                               **wkw)
        assert torch.isfinite(out.float()).all() and torch.isfinite(lse).all()
        # the second-generation kernels keep bias * log2e as bf16 in shared memory (one 16-byte vector per 8 keys): its rounding,
        # <= 2^-9 |bias log2e| on the exponent, is negligible for real tables (|bias| of a few units) and ~0.2 at this test's +-75
        assert rel_l2(out.view(B_, N, -1), oref) < (5e-2 if (hint and case == "huge_bias") else 3e-2)
        assert rel_l2(lse, lref) < 1e-2
        dout = rnd(B_, N, nH * hd, dtype=dtype)
        oref.backward(dout.double())
        dqkv, dtab = VF.attn_bwd(qkv.view(B_ * N, -1), out, dout.view(B_ * N, -1), lse, table, rowcode, colcode, plan.region,
                                 None, B_, nW, N, nH, hd, hd ** -0.5, planes=plan.ws[0] if hint else 0, **wkw)
        assert torch.isfinite(dqkv.float()).all() and torch.isfinite(dtab.float()).all()
        assert rel_l2(dqkv.view(B_, N, 3, nH, hd)[:, :, 2], qr.grad[:, :, 2]) < 6e-2   # dV (softmax is nearly one-hot here)
    finally:
        L.set_gemm_backend(0)


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("backend", [0, 1])
def test_window_attention_swin_l_384_window(vsw, oracle, dtype, backend):
    """Window 8x12x12 (N = 1152, Swin-L 384, BASELINE configs 3/4): beyond the tcgen05 kernels' N <= 448, so AUTO must fall
    back to the CUDA-core kernels (and a forced tcgen05 back end must refuse loudly)."""
    VF, L = vsw.functional, vsw._lib
    grid, window, shift, nH, hd, B = (8, 24, 24), (8, 12, 12), (0, 6, 6), 2, 32, 1
    plan = VF.window_plan(grid, window, shift, "cuda")
    nW, N = plan.nW, plan.N
    assert N == 1152 and plan.shifted
    B_ = B * nW
    torch.manual_seed(5)
    qkv = rnd(B_, N, 3, nH, hd, dtype=dtype)
    Lt = 15 * 23 * 23
    table = rnd(Lt, nH, dtype=dtype, scale=0.5)
    rel_index = torch.from_numpy(oracle.relative_position_index(window)).cuda()
    rowcode, colcode = VF.bias_codes(rel_index, N)
    mask = vsw.compute_mask(*plan.pgrid, plan.ws, plan.ss, "cuda")
    qr = qkv.double().requires_grad_(True)
    tr = table.double().requires_grad_(True)
    oref, lref = attn_reference(qr, tr, rel_index, mask, nW, nH, hd ** -0.5)
    L.set_gemm_backend(backend)
    try:
        out, lse = VF.attn_fwd(qkv.view(B_ * N, -1), table, rowcode, colcode, plan.region, None, B_, nW, N, nH, hd, hd ** -0.5)
        assert rel_l2(out.view(B_, N, -1), oref) < TOL[dtype]
        dout = rnd(B_, N, nH * hd, dtype=dtype)
        oref.backward(dout.double())
        dqkv, dtab = VF.attn_bwd(qkv.view(B_ * N, -1), out, dout.view(B_ * N, -1), lse, table, rowcode, colcode, plan.region,
                                 None, B_, nW, N, nH, hd, hd ** -0.5)
        assert rel_l2(dqkv.view(B_, N, 3, nH, hd), qr.grad) < TOL[dtype] * (1 if dtype == torch.float32 else 1.5)
        assert rel_l2(dtab, tr.grad) < TOL[dtype] * (1 if dtype == torch.float32 else 1.5)
        if dtype == torch.bfloat16:
            L.set_gemm_backend(L.GEMM_TCGEN05)
            with pytest.raises(L.VswError):
                VF.attn_fwd(qkv.view(B_ * N, -1), table, rowcode, colcode, plan.region, None, B_, nW, N, nH, hd, hd ** -0.5)
    finally:
        L.set_gemm_backend(0)


# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("shape", [(2, 3, 8, 32, 32), (1, 3, 4, 30, 27), (2, 3, 3, 8, 8)])
def test_patch_embed(vsw, dtype, shape):
    VF = vsw.functional
    B, Cin, D, H, W = shape
    E, patch = 64, (2, 4, 4)
    x = rnd(*shape, seed=3)  # clips arrive in fp32
    w, b = rnd(E, Cin, *patch, dtype=dtype, scale=0.1), rnd(E, dtype=dtype, scale=0.1)
    g, be = rnd(E, dtype=dtype) * 0.2 + 1, rnd(E, dtype=dtype) * 0.1
    xr = x.double().requires_grad_(True)
    wr, br = w.double().requires_grad_(True), b.double().requires_grad_(True)
    xp = F.pad(xr, (0, (-W) % 4, 0, (-H) % 4, 0, 1))
    yr = F.conv3d(xp, wr, br, stride=(1, 4, 4)).permute(0, 2, 3, 4, 1)
    yr = F.layer_norm(yr, (E,), g.double(), be.double(), 1e-5)
    xg = x.clone().requires_grad_(True)
    wg, bg = w.clone().requires_grad_(True), b.clone().requires_grad_(True)
    y = VF.patch_embed(xg, wg, bg, g, be, patch)
    assert rel_l2(y.view(yr.shape), yr) < TOL[dtype]
    dy = rnd(*yr.shape, dtype=dtype)
    yr.backward(dy.double())
    y.backward(dy.view(y.shape))
    assert rel_l2(wg.grad, wr.grad) < TOL[dtype]
    assert rel_l2(bg.grad, br.grad) < TOL[dtype]
    assert rel_l2(xg.grad, xr.grad) < TOL[dtype]


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("grid", [(2, 8, 8), (3, 7, 9)])
def test_patch_merge(vsw, dtype, grid):
    VF = vsw.functional
    D, H, W = grid
    B, C = 2, 32
    x = rnd(B, D, H, W, C, dtype=dtype, seed=9)
    g, be = rnd(4 * C, dtype=dtype) * 0.2 + 1, rnd(4 * C, dtype=dtype) * 0.1
    w = rnd(2 * C, 4 * C, dtype=dtype, scale=0.1)
    xr = x.double().requires_grad_(True)
    gr, ber, wr = [t.double().requires_grad_(True) for t in (g, be, w)]
    xp = F.pad(xr, (0, 0, 0, W % 2, 0, H % 2))
    cat = torch.cat([xp[:, :, 0::2, 0::2], xp[:, :, 1::2, 0::2], xp[:, :, 0::2, 1::2], xp[:, :, 1::2, 1::2]], -1)
    yr = F.layer_norm(cat, (4 * C,), gr, ber, 1e-5) @ wr.t()
    xs = [t.clone().requires_grad_(True) for t in (x, g, be, w)]
    y = VF.patch_merge(xs[0].view(B, D * H * W, C), xs[1], xs[2], xs[3], grid)
    assert rel_l2(y.view(yr.shape), yr) < TOL[dtype]
    dy = rnd(*yr.shape, dtype=dtype)
    yr.backward(dy.double())
    y.backward(dy.view(y.shape))
    for got, ref, nm in zip(xs, (xr, gr, ber, wr), ("x", "gamma", "beta", "w")):
        assert rel_l2(got.grad, ref.grad) < TOL[dtype], nm


# ---------------------------------------------------------------------------------------------
# EncVideo tail (model.py:57-76): class row + position / frame embeddings + LayerNorm + mask
# ---------------------------------------------------------------------------------------------
def _tail_params(hid, max_frame, max_patch, seed):
    torch.manual_seed(seed)
    return {"emb_cls": 0.3 * torch.randn(1, 1, 1, hid), "emb_pos": 0.3 * torch.randn(1, 1, 1 + max_patch ** 2, hid),
            "emb_len": 0.3 * torch.randn(1, max_frame, 1, hid), "emb_odr": 0.3 * torch.randn(1, 1, 1, hid),
            "norm.weight": 1 + 0.2 * torch.randn(hid), "norm.bias": 0.1 * torch.randn(hid)}


@pytest.mark.parametrize("dtype,out_dtype", [(torch.float32, None), (torch.bfloat16, None), (torch.bfloat16, torch.float32),
                                             (torch.float16, None)])
@pytest.mark.parametrize("B,Tn,h,w,hid,use_odr", [(2, 3, 2, 2, 24, False), (3, 4, 2, 3, 768, True), (5, 8, 7, 7, 768, True),
                                                  (1, 1, 1, 1, 1024, False), (2, 6, 3, 3, 132, True)])
def test_enc_video_tail_vs_oracle(vsw, dtype, out_dtype, B, Tn, h, w, hid, use_odr):
    from oracle import enc_video_oracle as EO
    VF = vsw.functional
    hw, P = h * w, 1 + h * w
    p = _tail_params(hid, max(Tn, 6), 14, seed=B * 100 + Tn)
    torch.manual_seed(hid + Tn)
    f = torch.randn(B, Tn, hw, hid).to(dtype)
    odr = torch.stack([torch.randperm(Tn) for _ in range(B)]) if use_odr else None
    vt = (torch.rand(B, Tn, P) > 0.25).long() if use_odr else None
    R = torch.randn(B, Tn * P, hid)
    # oracle in fp64 on the CPU, from the same (storage-rounded) features
    fr = f.double().requires_grad_(True)
    pr = {k: v.double().requires_grad_(True) for k, v in p.items()}
    # the oracle takes the backbone layout (B, L, T, h, w); no fc here
    o_ref, m_ref = EO.enc_video_tail(fr.view(B, Tn, h, w, hid).permute(0, 4, 1, 2, 3), pr, odr=odr, vt_mask=vt)
    (o_ref * R.double()).sum().backward()
    # CUDA path
    fc = f.cuda().requires_grad_(True)
    pc = {k: v.cuda().requires_grad_(True) for k, v in p.items()}
    out, m_img = VF.enc_video_tail(fc, pc["emb_cls"], pc["emb_pos"], pc["emb_len"], pc["emb_odr"], pc["norm.weight"],
                                   pc["norm.bias"], None if odr is None else odr.to(torch.int32).cuda(),
                                   None if vt is None else vt.cuda(), out_dtype)
    assert out.dtype == (out_dtype or dtype) and out.shape == (B, Tn * P, hid)
    assert m_img.dtype == torch.int64 and torch.equal(m_img.cpu(), m_ref)           # integer domain: bit-exact
    tol = TOL[out_dtype or dtype]
    assert rel_l2(out, o_ref) < tol
    (out.float() * R.cuda()).sum().backward()
    assert rel_l2(fc.grad, fr.grad) < TOL[dtype]
    for k in p:
        ref = pr[k].grad
        if ref is None or float(ref.abs().max()) == 0.0:       # emb_odr without odr; unused emb_pos / emb_len rows
            assert pc[k].grad is None or float(pc[k].grad.abs().max()) == 0.0, k
        else:
            # fp32 accumulation over dy in the storage dtype of `out`
            assert rel_l2(pc[k].grad, ref) < (1e-4 if (out_dtype or dtype) == torch.float32 else tol), k
    assert float(pc["emb_pos"].grad[0, 0, P:].abs().max() if P < pc["emb_pos"].shape[2] else 0.0) == 0.0   # unused rows: exactly 0
    assert float(pc["emb_len"].grad[0, Tn:].abs().max() if Tn < pc["emb_len"].shape[1] else 0.0) == 0.0


def test_enc_video_tail_golden_and_determinism(vsw):
    """the reference's own outputs (tests/golden/enc_video.pt), and bit-identical reruns (fixed reduction orders)"""
    import os
    VF = vsw.functional
    gold = torch.load(os.path.join(os.path.dirname(__file__), "golden", "enc_video.pt"), weights_only=False)
    for case in ("plain", "odr", "nofc_vtmask"):
        g = gold[case]
        prm = {k: v.cuda().requires_grad_(True) for k, v in g["params"].items()}
        buf = g["buf"].cuda().requires_grad_(True)
        B, Tn, h, w, L = buf.shape
        runs = []
        for _ in range(2):
            for t in [buf, *prm.values()]:
                t.grad = None
            tok = buf.view(B, Tn, h * w, L)
            if "fc.weight" in prm:
                tok = VF.linear(tok.view(-1, L), prm["fc.weight"], prm["fc.bias"]).view(B, Tn, h * w, -1)
            odr = None if g["odr"] is None else torch.tensor(g["odr"], dtype=torch.int32, device="cuda")
            vt = None if g["vt_mask"] is None else g["vt_mask"].cuda()
            out, m_img = VF.enc_video_tail(tok, prm["emb_cls"], prm["emb_pos"], prm["emb_len"], prm["emb_odr"],
                                           prm["norm.weight"], prm["norm.bias"], odr, vt, None)
            (out * g["R"].cuda()).sum().backward()
            runs.append([out.detach().clone(), buf.grad.clone()] + [prm[k].grad.clone() for k in sorted(g["grads"])])
        assert torch.equal(m_img.cpu(), g["m_img"])
        assert rel_l2(runs[0][0], g["f_img"]) < 1e-4
        assert rel_l2(runs[0][1], g["dbuf"]) < 1e-4
        for k, got in zip(sorted(g["grads"]), runs[0][2:]):
            assert rel_l2(got, g["grads"][k]) < 1e-4, (case, k)
        for a, b in zip(runs[0], runs[1]):
            assert torch.equal(a, b)


def test_enc_video_tail_rejects_bad_geometry(vsw):
    VF = vsw.functional
    p = {k: v.cuda() for k, v in _tail_params(24, 2, 1, seed=0).items()}
    args = (p["emb_cls"], p["emb_pos"], p["emb_len"], p["emb_odr"], p["norm.weight"], p["norm.bias"])
    with pytest.raises(vsw._lib.VswError):       # 3 frames, emb_len holds 2 (the reference's add fails to broadcast)
        VF.enc_video_tail(torch.zeros(1, 3, 1, 24, device="cuda"), *args)
    with pytest.raises(vsw._lib.VswError):       # 1 + 4 tokens per frame, emb_pos holds 2
        VF.enc_video_tail(torch.zeros(1, 2, 4, 24, device="cuda"), *args)
    with pytest.raises(vsw._lib.VswError):       # CPU tensor: no fallback
        VF.enc_video_tail(torch.zeros(1, 2, 1, 24), *args)


# ---------------------------------------------------------------------------------------------
# MVM masking / 3d_feature loss (main_pretrain.py:355-362, 508-524)
# ---------------------------------------------------------------------------------------------
def test_block_mask_apply_vs_reference_golden(vsw):
    """the reference's own masked clip (fp32: bit-exact), with the blocks re-drawn by the product sampler"""
    import os
    import numpy as np
    from conftest import pattern_clip
    g = torch.load(os.path.join(os.path.dirname(__file__), "golden", "mvm.pt"), weights_only=False)["masking"]
    B, Tn, H, W = g["shape"]
    np.random.seed(g["np_seed"])
    cov = vsw.mvm.sample_block_masks(B, Tn, H // 32, W // 32)
    img = pattern_clip(B, Tn, H, W).cuda()
    masked, mask = vsw.mvm.apply_block_mask(img, cov, 32)
    assert masked.data_ptr() != img.data_ptr() and torch.equal(img.cpu(), pattern_clip(B, Tn, H, W))   # not in place
    assert torch.equal(masked[:, :, :, ::37, :].cpu(), g["check_rows"])
    assert torch.equal(mask[:, :, :, ::37, :].to(torch.uint8).cpu(), g["mask_rows"])
    assert float(mask.sum()) == g["mask_sum"]
    assert torch.equal(torch.nn.functional.max_pool2d(mask.view(B * Tn, 3, H, W), 32).view(B, Tn, 3, 7, 7).to(torch.uint8).cpu(),
                       g["cover"])
    # in place, clip only
    img2 = img.clone()
    m2, none = vsw.mvm.apply_block_mask(img2, cov, 32, inplace=True, want_mask=False)
    assert none is None and m2.data_ptr() == img2.data_ptr() and torch.equal(img2, masked)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16, torch.float16])
@pytest.mark.parametrize("B,Tn,H,W,ps", [(2, 3, 64, 96, 32), (1, 2, 48, 80, 16), (3, 1, 32, 32, 32)])
def test_block_mask_apply_vs_oracle(vsw, dtype, B, Tn, H, W, ps):
    from oracle import mvm_oracle as MO
    torch.manual_seed(H + W)
    img = torch.randn(B, Tn, 3, H, W).to(dtype)
    cov = (torch.rand(B, Tn, H // ps, W // ps) > 0.5)
    ref, mref = MO.apply_block_mask(img, cov.float(), ps)
    out, mask = vsw.mvm.apply_block_mask(img.cuda(), cov.to(torch.uint8), ps)
    assert out.dtype == dtype and torch.equal(out.cpu(), ref) and torch.equal(mask.cpu(), mref)      # exact: x*1, x*0
    with pytest.raises(vsw._lib.VswError):
        vsw.mvm.apply_block_mask(img.cuda(), torch.zeros(B, Tn, H // ps + 1, W // ps, dtype=torch.uint8), ps)   # wrong grid


@pytest.mark.parametrize("dtype,tdtype", [(torch.float32, torch.float32), (torch.bfloat16, torch.bfloat16),
                                          (torch.bfloat16, torch.float32), (torch.float16, torch.float16)])
@pytest.mark.parametrize("rows,C,frac", [(37, 16, 0.5), (12544, 1024, 0.15), (300, 132, 1.0), (64, 768, 0.0)])
def test_masked_l1_vs_oracle(vsw, dtype, tdtype, rows, C, frac):
    VF = vsw.functional
    torch.manual_seed(rows + C)
    pred = torch.randn(rows, C).to(dtype)
    target = torch.randn(rows, C).to(tdtype)
    target[::3, ::5] = pred[::3, ::5].to(tdtype)                      # exact ties: sign(0) = 0
    m = (torch.rand(rows) < frac).float()
    pr = pred.double().requires_grad_(True)
    ref = (torch.nn.functional.l1_loss(pr, target.double(), reduction="none") * m.double().view(-1, 1)).sum() / (m.double().sum() + 1e-5) / 3
    ref.backward()
    pc = pred.cuda().requires_grad_(True)
    runs = []
    for _ in range(2):
        pc.grad = None
        loss = VF.masked_l1(pc, target.cuda(), m.cuda(), 3.0)
        (loss * 2.5).backward()
        runs.append((loss.detach().clone(), pc.grad.clone()))
    loss, grad = runs[0]
    assert loss.dtype == torch.float32 and loss.ndim == 0
    assert abs(float(loss) - float(ref)) <= 2e-5 * abs(float(ref)) + 1e-12
    gref = pr.grad * 2.5
    if float(gref.abs().max()) == 0.0:
        assert float(grad.abs().max()) == 0.0
    else:
        assert rel_l2(grad, gref) < (1e-5 if dtype == torch.float32 else TOL[dtype])
        assert torch.equal(grad == 0, (gref == 0).cuda())             # masked-out rows and exact ties carry no gradient
    assert torch.equal(runs[0][0], runs[1][0]) and torch.equal(runs[0][1], runs[1][1])   # fixed reduction order


def test_mvm_3d_feature_loss_vs_reference_golden(vsw):
    """main_pretrain.py:508-524 end to end through the C ABI: fc_mvm (vsw_linear) on the non-class rows, masked L1 against
    the teacher tokens read from a Swin-style permuted view; loss and gradients vs the reference's own values"""
    import os
    g = torch.load(os.path.join(os.path.dirname(__file__), "golden", "mvm.pt"), weights_only=False)["loss"]
    VF = vsw.functional
    B, Tn, h, w, Cf = g["teacher"].shape
    P = 1 + h * w
    out_mvm = g["out_mvm"].cuda().requires_grad_(True)
    fw, fb = g["fc_w"].cuda().requires_grad_(True), g["fc_b"].cuda().requires_grad_(True)
    non_cls = out_mvm.view(B, Tn, P, -1)[:, :, 1:].reshape(B * Tn * h * w, -1)
    pred = VF.linear(non_cls, fw, fb).view(B, Tn, h * w, Cf)
    teacher_view = g["teacher"].cuda().permute(0, 4, 1, 2, 3)          # (B,C,T,h,w) view of the channels-last buffer
    loss = vsw.mvm.mvm_3d_feature_loss(pred, teacher_view, g["cov"], 3)
    assert abs(float(loss) - float(g["loss"])) < 1e-5 * abs(float(g["loss"]))
    loss.backward()
    assert rel_l2(out_mvm.grad, g["d_out_mvm"]) < 1e-4
    assert rel_l2(fw.grad, g["d_fc_w"]) < 1e-4 and rel_l2(fb.grad, g["d_fc_b"]) < 1e-4


@pytest.mark.parametrize("xdtype,ydtype", [(torch.float32, torch.bfloat16), (torch.float32, torch.float16),
                                           (torch.bfloat16, torch.bfloat16)])
@pytest.mark.parametrize("mapped", [False, True])
def test_residual_add(vsw, xdtype, ydtype, mapped):
    """vsw_residual_add: out[b, map[r]] = x[b, map[r]] + scale[b] * y[b, r] in the (wider) dtype of x -- the fp32 residual
    stream of torch.autocast (video_swin.py:256, 261 promote `shortcut + drop_path(x)` to fp32 there)."""
    VF = vsw.functional
    B, C = 3, 64
    plan = VF.window_plan((8, 10, 10), (8, 7, 7), (0, 3, 3), "cuda") if mapped else None
    T = 800
    R = plan.nW * plan.N if mapped else T
    x = rnd(B, T, C, dtype=xdtype)
    y = rnd(B, R, C, dtype=ydtype)
    scale = torch.tensor([0.0, 1.25, 1.25], device="cuda")
    out = VF.residual_add(x, y, plan.gather if mapped else None, scale, B, R, T, C)
    ref = x.double().clone()
    if mapped:
        g = plan.gather.long()
        ok = g >= 0
        ref[:, g[ok]] += scale.double()[:, None, None] * y.double()[:, ok]
    else:
        ref += scale.double()[:, None, None] * y.double()
    assert out.dtype == xdtype
    assert rel_l2(out, ref) < (1e-6 if xdtype == torch.float32 else 5e-3)


@pytest.mark.parametrize("dtype", DTYPES)
def test_drop_path_scale_kernel(vsw, dtype):
    """one launch instead of add / floor / cast / div; the sum keep + u is rounded to the dtype of u as torch rounds it"""
    L = vsw._lib
    torch.manual_seed(3)
    u = torch.rand(4096, device="cuda", dtype=dtype)
    for keep in (0.8, 0.9130434782608696, 0.5):
        out = torch.empty(4096, device="cuda")
        L.check(L.lib().vsw_drop_path_scale(L.ptr(u), float(keep), L.ptr(out), 4096, L.dt(u), L.stream()), "drop_path_scale")
        ref = (keep + u).floor().float() / keep
        assert torch.equal(out, ref), keep
