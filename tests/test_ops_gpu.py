"""GPU parity of the individual kernels (through the C ABI) against plain fp32/fp64 torch references."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

from bevgen_b200 import ops  # noqa: E402


def dev():
    return torch.device("cuda:0")


def _split(x):
    hi = x.to(torch.bfloat16)
    lo = (x - hi.float()).to(torch.bfloat16)
    return hi.contiguous(), lo.contiguous()


def _ref_planes(hi, lo, npass):
    return (hi.double() + lo.double()) if npass == 3 else hi.double()


@pytest.mark.parametrize("npass", [3, 1])
@pytest.mark.parametrize("bn", [128, 64, 16])
@pytest.mark.parametrize("M,N,K", [(256, 128, 64), (1000, 200, 320), (128, 3, 128)])
def test_linear(M, N, K, bn, npass):
    g = torch.Generator().manual_seed(M + N + K)
    a = torch.randn(M, K, generator=g).to(dev())
    w = (torch.randn(N, K, generator=g) / K ** 0.5).to(dev())
    bias = torch.randn(N, generator=g).to(dev())
    res = torch.randn(M, N, generator=g).to(dev())
    a_hi, a_lo = _split(a)
    rows_pad = max(bn, ((N + 15) // 16) * 16)
    wp = torch.zeros(rows_pad, K, device=dev())
    wp[:N] = w
    w_hi, w_lo = _split(wp)
    out = torch.full((M, N), float("nan"), device=dev())
    ops.gemm_tc(a_hi=a_hi, a_lo=a_lo, a_dims=(1, 1, M, K), b_hi=w_hi, b_lo=w_lo, k=K, n_cols=N, out_w=M, ldc=N,
                bias=bias, residual=res, out_f32=out, bn=bn, npass=npass)
    torch.cuda.synchronize()
    if npass == 3:
        ref = a.double() @ w.double().t() + bias.double() + res.double()
        tol = 2e-5 * K ** 0.5
    else:
        ref = a_hi.double() @ w_hi[:N].double().t() + bias.double() + res.double()
        tol = 1e-5 * K ** 0.5
    err = (out.double() - ref).abs().max().item()
    assert torch.isfinite(out).all()
    assert err < tol, f"max err {err}"


@pytest.mark.parametrize("npass", [3, 1])
def test_linear_split_out_gelu(npass):
    M, N, K = 384, 256, 128
    g = torch.Generator().manual_seed(5)
    a = torch.randn(M, K, generator=g).to(dev())
    w = (torch.randn(N, K, generator=g) / K ** 0.5).to(dev())
    bias = torch.randn(N, generator=g).to(dev())
    a_hi, a_lo = _split(a)
    w_hi, w_lo = _split(w)
    o_hi = torch.zeros(M, N, dtype=torch.bfloat16, device=dev())
    o_lo = torch.zeros(M, N, dtype=torch.bfloat16, device=dev())
    ops.gemm_tc(a_hi=a_hi, a_lo=a_lo, a_dims=(1, 1, M, K), b_hi=w_hi, b_lo=w_lo, k=K, n_cols=N, out_w=M, ldc=N, bias=bias,
                out_hi=o_hi, out_lo=o_lo, flags=ops.GF_GELU, bn=128, npass=npass)
    torch.cuda.synchronize()
    ref = F.gelu(_ref_planes(a_hi, a_lo, npass) @ _ref_planes(w_hi, w_lo, npass).t() + bias.double())
    got = o_hi.double() + o_lo.double()
    assert (got - ref).abs().max().item() < 5e-5


@pytest.mark.parametrize("npass", [3, 1])
@pytest.mark.parametrize("N_,H,W,Cin,Cout,bn", [(2, 16, 16, 64, 128, 128), (1, 32, 32, 128, 64, 64), (3, 8, 8, 128, 3, 16),
                                                 (2, 4, 4, 64, 64, 64), (1, 24, 40, 64, 192, 128)])
def test_conv3x3(N_, H, W, Cin, Cout, bn, npass):
    g = torch.Generator().manual_seed(H * W + Cin)
    x = torch.randn(N_, Cin, H, W, generator=g).to(dev())
    w = (torch.randn(Cout, Cin, 3, 3, generator=g) / (9 * Cin) ** 0.5).to(dev())
    b = torch.randn(Cout, generator=g).to(dev())
    x_nhwc = x.permute(0, 2, 3, 1).contiguous()
    a_hi, a_lo = _split(x_nhwc)
    rows = 9 * Cout
    rows_pad = max(rows, 8 * Cout + bn)
    wp = torch.zeros(rows_pad, Cin, device=dev())
    wp[:rows] = w.permute(2, 3, 0, 1).reshape(rows, Cin)      # [tap=(kh,kw)][cout][cin]
    w_hi, w_lo = _split(wp)
    nchw = Cout < 8
    out = torch.full((N_, Cout, H, W) if nchw else (N_, H, W, Cout), float("nan"), device=dev())
    tw = 16 if W >= 16 else 8
    ops.gemm_tc(a_hi=a_hi, a_lo=a_lo, a_dims=(N_, H, W, Cin), b_hi=w_hi, b_lo=w_lo, k=Cin, n_cols=Cout,
                taps=[(dx, dy, 0) for dx, dy in ops.TAPS_3X3], b_row_tapstride=Cout, z_outer=N_, tile=(tw, 128 // tw),
                out_w=W, out_h=H, out_zo_stride=H * W * Cout, ldc=Cout, bias=b, out_f32=out,
                flags=ops.GF_OUT_NCHW if nchw else 0, bn=bn, npass=npass)
    torch.cuda.synchronize()
    xin = _ref_planes(a_hi, a_lo, npass).permute(0, 3, 1, 2)
    win = _ref_planes(w_hi, w_lo, npass)[:rows].reshape(3, 3, Cout, Cin).permute(2, 3, 0, 1)
    ref = F.conv2d(xin, win, b.double(), padding=1)
    got = out.double() if nchw else out.double().permute(0, 3, 1, 2)
    assert torch.isfinite(out).all()
    err = (got - ref).abs().max().item()
    assert err < (3e-5 if npass == 3 else 1e-5) * 10, f"max err {err}"


def test_conv_stride2_space_to_depth():
    """Downsample (model.py:68-75): pad (0,1,0,1) + 3x3 stride 2 == 9 taps over 4 phase planes."""
    N_, H, W, Cc = 2, 16, 16, 64
    g = torch.Generator().manual_seed(9)
    x = torch.randn(N_, Cc, H, W, generator=g).to(dev())
    w = (torch.randn(Cc, Cc, 3, 3, generator=g) / (9 * Cc) ** 0.5).to(dev())
    b = torch.randn(Cc, generator=g).to(dev())
    x_nhwc = x.permute(0, 2, 3, 1).contiguous()
    hi = torch.empty(N_ * 4, H // 2, W // 2, Cc, dtype=torch.bfloat16, device=dev())
    lo = torch.empty_like(hi)
    ops.prep_operand(x_nhwc, hi, lo, mode=ops.PREP_S2D)
    wp = torch.zeros(9 * Cc + 64, Cc, device=dev())
    wp[:9 * Cc] = w.permute(2, 3, 0, 1).reshape(9 * Cc, Cc)
    w_hi, w_lo = _split(wp)
    out = torch.full((N_, H // 2, W // 2, Cc), float("nan"), device=dev())
    taps = [(kw // 2, kh // 2, (kh & 1) * 2 + (kw & 1)) for kh in range(3) for kw in range(3)]
    ops.gemm_tc(a_hi=hi, a_lo=lo, a_dims=(N_ * 4, H // 2, W // 2, Cc), b_hi=w_hi, b_lo=w_lo, k=Cc, n_cols=Cc, taps=taps, a_n_mul=4,
                b_row_tapstride=Cc, z_outer=N_, tile=(8, 16), out_w=W // 2, out_h=H // 2, out_zo_stride=(H // 2) * (W // 2) * Cc,
                ldc=Cc, bias=b, out_f32=out, bn=64, npass=3)
    torch.cuda.synchronize()
    ref = F.conv2d(F.pad(x.double(), (0, 1, 0, 1)), w.double(), b.double(), stride=2)
    err = (out.double().permute(0, 3, 1, 2) - ref).abs().max().item()
    assert err < 1e-4, f"max err {err}"


def test_batched_qk_and_pv_mn_major():
    """S = Q K^T per (batch, head) from a fused qkv buffer, then O = P V with V read MN-major (no transpose)."""
    B, Hh, L, dh = 2, 4, 256, 64
    d = Hh * dh
    g = torch.Generator().manual_seed(3)
    qkv = torch.randn(B, L, 3 * d, generator=g).to(dev())
    q_hi, q_lo = _split(qkv)
    S = torch.full((B, Hh, L, L), float("nan"), device=dev())
    q2h, q2l = q_hi.view(B * L, 3 * d), q_lo.view(B * L, 3 * d)
    ops.gemm_tc(a_hi=q_hi, a_lo=q_lo, a_dims=(B, 1, L, 3 * d), b_hi=q2h, b_lo=q2l, k=dh, n_cols=L, a_c_off=0, a_c_zstride=dh,
                b_k_off=d, b_k_zstride=dh, b_row_zstride=L, z_inner=Hh, z_outer=B, out_w=L, out_zo_stride=Hh * L * L,
                out_zi_stride=L * L, ldc=L, out_f32=S, bn=128, npass=3)
    torch.cuda.synchronize()
    qd = qkv.double().view(B, L, 3, Hh, dh)
    ref_S = torch.einsum("blhd,bmhd->bhlm", qd[:, :, 0], qd[:, :, 1])
    err = (S.double() - ref_S).abs().max().item()
    assert err < 2e-5 * ref_S.abs().max().item(), f"QK^T max err {err}"
    P = torch.softmax(ref_S / 8, -1).float()
    p_hi, p_lo = _split(P)
    O = torch.full((B, L, d), float("nan"), device=dev())
    ops.gemm_tc(a_hi=p_hi, a_lo=p_lo, a_dims=(B * Hh, 1, L, L), b_hi=q2h, b_lo=q2l, k=L, n_cols=dh, a_n_mul=Hh, a_n_zstride=1,
                b_k_off=2 * d, b_k_zstride=dh, b_row_zstride=L, z_inner=Hh, z_outer=B, out_w=L, out_zo_stride=L * d,
                out_zi_stride=dh, ldc=d, out_f32=O, flags=ops.GF_B_MN, bn=64, npass=3)
    torch.cuda.synchronize()
    ref_O = torch.einsum("bhlm,bmhd->blhd", P.double(), qd[:, :, 2]).reshape(B, L, d)
    err = (O.double() - ref_O).abs().max().item()
    assert err < 2e-5, f"PV max err {err}"


@pytest.mark.parametrize("C_", [64, 128, 256, 512])
def test_groupnorm_prep(C_):
    N_, H, W = 3, 16, 8
    g = torch.Generator().manual_seed(C_)
    x = (torch.randn(N_, C_, H, W, generator=g) * 2 + 0.5).to(dev())
    gamma = torch.randn(C_, generator=g).to(dev())
    beta = torch.randn(C_, generator=g).to(dev())
    x_nhwc = x.permute(0, 2, 3, 1).contiguous()
    ws = torch.empty(N_ * 64, dtype=torch.float64, device=dev())
    mr = torch.empty(N_ * 64, dtype=torch.float32, device=dev())
    ops.groupnorm_stats(x_nhwc, ws, mr, 1e-6)
    for swish, mode in [(True, ops.PREP_IDENT), (False, ops.PREP_IDENT), (True, ops.PREP_UP2)]:
        oh, ow = (2 * H, 2 * W) if mode == ops.PREP_UP2 else (H, W)
        hi = torch.empty(N_, oh, ow, C_, dtype=torch.bfloat16, device=dev())
        lo = torch.empty_like(hi)
        ops.prep_operand(x_nhwc, hi, lo, mr, gamma, beta, swish=swish, mode=mode)
        torch.cuda.synchronize()
        ref = F.group_norm(x.double(), 32, gamma.double(), beta.double(), eps=1e-6)
        if swish:
            ref = ref * torch.sigmoid(ref)
        if mode == ops.PREP_UP2:
            ref = F.interpolate(ref, scale_factor=2.0, mode="nearest")
        got = (hi.double() + lo.double()).permute(0, 3, 1, 2)
        err = ((got - ref).abs() / (1 + ref.abs())).max().item()
        assert err < 2e-5, f"C={C_} swish={swish} mode={mode}: {err}"


@pytest.mark.parametrize("cin", [3, 7])
def test_im2col(cin):
    N_, H, W = 2, 12, 20
    g = torch.Generator().manual_seed(cin)
    x = torch.randn(N_, cin, H, W, generator=g).to(dev())
    hi = torch.empty(N_, H, W, 64, dtype=torch.bfloat16, device=dev())
    lo = torch.empty_like(hi)
    ops.im2col3x3(x, hi, lo)
    torch.cuda.synchronize()
    cols = F.unfold(x, 3, padding=1).view(N_, cin, 9, H, W).permute(0, 3, 4, 2, 1).reshape(N_, H, W, 9 * cin)
    got = hi.float() + lo.float()
    assert (got[..., :9 * cin] - cols).abs().max().item() < 1e-4
    assert (got[..., 9 * cin:] == 0).all()


@pytest.mark.parametrize("cb", ["normal", "default"])
def test_vq_nearest_golden(cb, golden_dir):
    """Bit-exact indices vs the reference's VectorQuantizer2 on the committed golden (separated + default codebooks)."""
    from oracle import synth
    gold = np.load(golden_dir / f"vq_{cb}.npz")
    sd = synth.vqgan_state_dict(synth.vqgan_ddconfig(), seed=3, codebook=cb)
    z = synth.tensor_for("vq.z", (4, 256, 8, 8), seed=5, kind="embedding")
    if cb == "default":
        z = z * 1e-3
    book = sd["quantize.embedding.weight"].to(dev())
    zf = z.permute(0, 2, 3, 1).reshape(-1, 256).contiguous().to(dev())
    ee = torch.empty(1024, device=dev())
    ops.row_sqnorm(book, ee)
    idx = torch.empty(zf.shape[0], dtype=torch.int64, device=dev())
    zq = torch.empty_like(zf)
    ws = torch.empty(zf.shape[0], device=dev())
    ops.vq_nearest(zf, book, ee, ws, idx, zq)
    torch.cuda.synchronize()
    if cb == "normal":
        assert np.array_equal(idx.cpu().numpy(), gold["idx"])
    else:
        # degenerate near-tie codebook: the chosen code must be within a few ulps of the true minimum
        d = torch.cdist(zf.double(), book.double()) ** 2
        chosen = d.gather(1, idx[:, None])[:, 0]
        assert ((chosen - d.min(1).values) <= 1e-9 + 4e-7 * (zf.double() ** 2).sum(1)).all()
        assert (idx.cpu().numpy() == gold["idx"]).mean() > 0.95
    assert torch.equal(zq, book[idx])


def test_vq_large_and_ragged():
    g = torch.Generator().manual_seed(0)
    for rows in (1, 63, 24576):
        z = torch.randn(rows, 256, generator=g).to(dev())
        book = torch.randn(1024, 256, generator=g).to(dev())
        ee = torch.empty(1024, device=dev())
        ops.row_sqnorm(book, ee)
        idx = torch.empty(rows, dtype=torch.int64, device=dev())
        ws = torch.empty(rows, device=dev())
        ops.vq_nearest(z, book, ee, ws, idx, None)
        ref = torch.cdist(z.double(), book.double()).argmin(1)
        assert torch.equal(idx, ref)


def test_softmax_transpose_gather_denorm():
    g = torch.Generator().manual_seed(1)
    s = torch.randn(700, 256, generator=g).to(dev()) * 5
    hi = torch.empty(700, 256, dtype=torch.bfloat16, device=dev())
    lo = torch.empty_like(hi)
    ops.softmax_rows(s, hi, lo, 0.125)
    ref = torch.softmax(s.double() * 0.125, -1)
    assert ((hi.double() + lo.double()) - ref).abs().max().item() < 1e-6
    src = torch.randn(3, 50, 70, generator=g).to(dev())
    dst = torch.empty(3, 70, 50, device=dev())
    ops.transpose_f32(src, dst, 3, 50, 70)
    assert torch.equal(dst, src.transpose(1, 2).contiguous())
    book = torch.randn(1024, 256, generator=g).to(dev())
    idx = torch.randint(0, 1024, (513,), generator=g).to(dev())
    out = torch.empty(513, 256, device=dev())
    ops.codebook_gather(book, idx, out)
    assert torch.equal(out, book[idx])
    x = torch.randn(2, 3, 8, 8, generator=g).to(dev()) * 3
    o = torch.empty_like(x)
    mean, std = [0.4265, 0.4489, 0.4769], [0.2053, 0.2206, 0.2578]
    ops.denormalize(x, o, mean, std)
    ref = torch.clamp(x * torch.tensor(std, device=dev()).view(1, 3, 1, 1) + torch.tensor(mean, device=dev()).view(1, 3, 1, 1), 0, 1)
    assert (o - ref).abs().max().item() < 1e-6


@pytest.mark.parametrize("npass", [3, 1])
@pytest.mark.parametrize("N_,H,W,Cin,Cout", [(2, 16, 16, 64, 128), (1, 20, 28, 128, 128), (3, 8, 8, 128, 256), (2, 4, 4, 64, 64), (1, 64, 64, 128, 128)])
def test_conv3x3_halo(N_, H, W, Cin, Cout, npass):
    """Halo-tile conv kernel (no-swizzle shifted A descriptors) + fused GroupNorm statistics of the output."""
    g = torch.Generator().manual_seed(H * W + Cin + Cout)
    x = torch.randn(N_, Cin, H, W, generator=g).to(dev())
    w = (torch.randn(Cout, Cin, 3, 3, generator=g) / (9 * Cin) ** 0.5).to(dev())
    b = torch.randn(Cout, generator=g).to(dev())
    res = torch.randn(N_, H, W, Cout, generator=g).to(dev())
    a_hi, a_lo = _split(x.permute(0, 2, 3, 1).contiguous())
    rows = 9 * Cout
    wp = torch.zeros(8 * Cout + ((Cout + 127) // 128) * 128, Cin, device=dev())
    wp[:rows] = w.permute(2, 3, 0, 1).reshape(rows, Cin)
    w_hi, w_lo = _split(wp)
    out = torch.full((N_, H, W, Cout), float("nan"), device=dev())
    sums = torch.full((N_ * 64,), float("nan"), dtype=torch.float64, device=dev()) if Cout >= 128 else None
    ops.conv3x3_halo(a_hi, a_lo if npass == 3 else None, (N_, H, W, Cin), w_hi, w_lo if npass == 3 else None, Cout, b, out, residual=res,
                     gn_sums=sums, npass=npass)
    torch.cuda.synchronize()
    xin = _ref_planes(a_hi, a_lo, npass).permute(0, 3, 1, 2)
    win = _ref_planes(w_hi, w_lo, npass)[:rows].reshape(3, 3, Cout, Cin).permute(2, 3, 0, 1)
    ref = F.conv2d(xin, win, b.double(), padding=1) + res.double().permute(0, 3, 1, 2)
    assert torch.isfinite(out).all()
    err = (out.double().permute(0, 3, 1, 2) - ref).abs().max().item()
    assert err < 3e-4, f"max err {err}"
    if sums is None:
        return
    # GroupNorm statistics of the output: sum and sum of squares per (image, group)
    o = out.double().permute(0, 3, 1, 2).reshape(N_, 32, -1)
    want = torch.stack([o.sum(-1), (o * o).sum(-1)], -1).reshape(-1)
    rel = ((sums - want).abs() / (1 + want.abs())).max().item()
    assert rel < 1e-5, f"gn sums rel err {rel}"
    mr = torch.empty(N_ * 64, dtype=torch.float32, device=dev())
    ops.groupnorm_finalize(sums, mr, N_, H * W, Cout, 1e-6)
    mean = o.mean(-1).reshape(-1)
    assert (mr.view(-1, 2)[:, 0].double() - mean).abs().max().item() < 1e-5


@pytest.mark.parametrize("npass", [3, 1])
@pytest.mark.parametrize("N_,H,W,Cin,Cout,mode", [(2, 16, 16, 64, 128, "gn_swish"), (1, 20, 28, 128, 128, "gn_swish"), (3, 8, 8, 128, 256, "plain"),
                                                   (2, 16, 16, 128, 128, "up2"), (1, 64, 64, 128, 128, "gn_swish")])
@pytest.mark.parametrize("two_cta", [False, True, "block16"])
@pytest.mark.timeout(400)
def test_conv3x3_fused_prologue(N_, H, W, Cin, Cout, mode, npass, two_cta):
    """Conv reading fp32 activations directly with GroupNorm-apply + swish (+ nearest 2x upsample) fused into the operand path."""
    g = torch.Generator().manual_seed(H * W + Cin + Cout + len(mode))
    hs, ws_ = (H // 2, W // 2) if mode == "up2" else (H, W)
    x = (torch.randn(N_, Cin, hs, ws_, generator=g) * 1.5 + 0.3).to(dev())
    w = (torch.randn(Cout, Cin, 3, 3, generator=g) / (9 * Cin) ** 0.5).to(dev())
    b = torch.randn(Cout, generator=g).to(dev())
    gamma, beta = torch.randn(Cin, generator=g).to(dev()), torch.randn(Cin, generator=g).to(dev())
    res = torch.randn(N_, H, W, Cout, generator=g).to(dev())
    x_nhwc = x.permute(0, 2, 3, 1).contiguous()
    rows = 9 * Cout
    wp = torch.zeros(8 * Cout + ((Cout + 127) // 128) * 128, Cin, device=dev())
    wp[:rows] = w.permute(2, 3, 0, 1).reshape(rows, Cin)
    w_hi, w_lo = _split(wp)
    affine = None
    ref_in = x.double()
    if mode == "gn_swish":
        sums = torch.empty(N_ * 64, dtype=torch.float64, device=dev())
        mr = torch.empty(N_ * 64, dtype=torch.float32, device=dev())
        ops.groupnorm_stats(x_nhwc, sums, mr, 1e-6)
        affine = torch.empty(N_, Cin, 2, device=dev())
        ops.groupnorm_affine(sums, gamma, beta, affine, N_, hs * ws_, Cin, 1e-6)
        ref_in = F.group_norm(ref_in, 32, gamma.double(), beta.double(), eps=1e-6)
        ref_in = ref_in * torch.sigmoid(ref_in)
    if mode == "up2":
        ref_in = F.interpolate(ref_in, scale_factor=2.0, mode="nearest")
    out = torch.full((N_, H, W, Cout), float("nan"), device=dev())
    osums = torch.full((N_ * 64,), float("nan"), dtype=torch.float64, device=dev())
    ops.conv3x3_fused(x_nhwc, w_hi, w_lo if npass == 3 else None, Cout, b, out, affine=affine, swish=(mode == "gn_swish"), up2=(mode == "up2"),
                      residual=res, gn_sums=osums, npass=npass, two_cta=(two_cta is True), block16=(two_cta == "block16"))
    torch.cuda.synchronize()
    ref = F.conv2d(ref_in, w.double(), b.double(), padding=1) + res.double().permute(0, 3, 1, 2)
    assert torch.isfinite(out).all()
    err = (out.double().permute(0, 3, 1, 2) - ref).abs().max().item()
    assert err < (3e-4 if npass == 3 else 8e-2), f"max err {err}"
    o = out.double().permute(0, 3, 1, 2).reshape(N_, 32, -1)
    want = torch.stack([o.sum(-1), (o * o).sum(-1)], -1).reshape(-1)
    assert ((osums - want).abs() / (1 + want.abs())).max().item() < 1e-5


@pytest.mark.parametrize("N_,H,W,Cin,Cout,mode", [(2, 16, 16, 64, 128, "gn_swish"), (1, 20, 28, 128, 128, "gn_swish"), (3, 8, 8, 128, 256, "plain"),
                                                   (2, 16, 16, 128, 128, "up2"), (1, 64, 64, 128, 128, "gn_swish"), (1, 24, 8, 256, 96, "big")])
@pytest.mark.parametrize("block16", [False, True])
@pytest.mark.timeout(400)
def test_conv3x3_fused_f16f8(N_, H, W, Cin, Cout, mode, block16):
    """fp16 + 2 x e4m3 split product (bevgen_conv3x3_fused_f16f8): same contract as the bf16x3 kernel, error bound ~2^-15 per product.
    "big" feeds activations far outside the e4m3 window (|x| up to ~400): the kernel must degrade to fp16 accuracy, not overflow."""
    g = torch.Generator().manual_seed(H * W + Cin + Cout + len(mode))
    hs, ws_ = (H // 2, W // 2) if mode == "up2" else (H, W)
    amp = 100.0 if mode == "big" else 1.5
    x = (torch.randn(N_, Cin, hs, ws_, generator=g) * amp + 0.3).to(dev())
    w = (torch.randn(Cout, Cin, 3, 3, generator=g) / (9 * Cin) ** 0.5).to(dev())
    b = torch.randn(Cout, generator=g).to(dev())
    gamma, beta = torch.randn(Cin, generator=g).to(dev()), torch.randn(Cin, generator=g).to(dev())
    res = torch.randn(N_, H, W, Cout, generator=g).to(dev())
    x_nhwc = x.permute(0, 2, 3, 1).contiguous()
    rows = 9 * Cout
    wp = torch.zeros(8 * Cout + ((Cout + 127) // 128) * 128, Cin, device=dev())
    wp[:rows] = w.permute(2, 3, 0, 1).reshape(rows, Cin)
    w16, w8pair, lo_scale = (ops.pack_f16f8_block if block16 else ops.pack_f16f8)(wp)
    assert w16.dtype == torch.float16 and w8pair.dtype == torch.uint8 and w8pair.shape == (wp.shape[0], 2 * Cin)
    affine = None
    ref_in = x.double()
    if mode == "gn_swish":
        sums = torch.empty(N_ * 64, dtype=torch.float64, device=dev())
        mr = torch.empty(N_ * 64, dtype=torch.float32, device=dev())
        ops.groupnorm_stats(x_nhwc, sums, mr, 1e-6)
        affine = torch.empty(N_, Cin, 2, device=dev())
        ops.groupnorm_affine(sums, gamma, beta, affine, N_, hs * ws_, Cin, 1e-6)
        ref_in = F.group_norm(ref_in, 32, gamma.double(), beta.double(), eps=1e-6)
        ref_in = ref_in * torch.sigmoid(ref_in)
    if mode == "up2":
        ref_in = F.interpolate(ref_in, scale_factor=2.0, mode="nearest")
    out = torch.full((N_, H, W, Cout), float("nan"), device=dev())
    osums = torch.full((N_ * 64,), float("nan"), dtype=torch.float64, device=dev()) if Cout >= 128 else None
    ops.conv3x3_fused_f16f8(x_nhwc, w16, w8pair, lo_scale, Cout, b, out, affine=affine, swish=(mode == "gn_swish"), up2=(mode == "up2"),
                            residual=res, gn_sums=osums, block16=block16)
    torch.cuda.synchronize()
    ref = F.conv2d(ref_in, w.double(), b.double(), padding=1) + res.double().permute(0, 3, 1, 2)
    assert torch.isfinite(out).all()
    err = (out.double().permute(0, 3, 1, 2) - ref).abs().max().item()
    scale = ref.abs().max().item()
    # fp16-only accuracy would be ~5e-4 * scale-ish; the split product is two orders better
    assert err < (1e-3 * scale if mode == "big" else 3e-4), f"max err {err} (output scale {scale})"
    if osums is not None:
        o = out.double().permute(0, 3, 1, 2).reshape(N_, 32, -1)
        want = torch.stack([o.sum(-1), (o * o).sum(-1)], -1).reshape(-1)
        assert ((osums - want).abs() / (1 + want.abs())).max().item() < 1e-5


@pytest.mark.parametrize("M,N,K", [(256, 128, 64), (1000, 256, 320), (384, 512, 1024)])
def test_linear_f16f8(M, N, K):
    """gemm_tc npass = 2: fp16 + 2 x e4m3 split product with host-packed operand planes; error bound ~2^-15 per product."""
    g = torch.Generator().manual_seed(M + N + K)
    a = torch.randn(M, K, generator=g).to(dev())
    w = (torch.randn(N, K, generator=g) / K ** 0.5).to(dev())
    bias = torch.randn(N, generator=g).to(dev())
    res = torch.randn(M, N, generator=g).to(dev())
    a16, apair = ops.pack_act_f16f8(a)
    w16, wpair, lo_scale = ops.pack_f16f8(w)
    out = torch.full((M, N), float("nan"), device=dev())
    ops.gemm_tc(a_hi=a16, a_lo=apair, a_dims=(1, 1, M, K), b_hi=w16, b_lo=wpair, k=K, n_cols=N, out_w=M, ldc=N, bias=bias, residual=res,
                out_f32=out, bn=128, npass=2, lo_scale=lo_scale)
    torch.cuda.synchronize()
    ref = a.double() @ w.double().t() + bias.double() + res.double()
    err = (out.double() - ref).abs().max().item()
    assert torch.isfinite(out).all()
    assert err < 6e-5 * K ** 0.5, f"max err {err}"


def test_f16f8_plane_producers_chain():
    """LayerNorm -> f16f8 planes -> GEMM + GELU -> f16f8 planes (epilogue) -> GEMM: the operand planes written by the kernels decode back to
    the fp32 values (fp16 + e4m3 remainder / 2^13) and the chain matches an fp64 reference."""
    M, D = 320, 256
    g = torch.Generator().manual_seed(3)
    x = (torch.randn(M, D, generator=g) * 2 + 0.5).to(dev())
    gamma, beta = (1 + 0.1 * torch.randn(D, generator=g)).to(dev()), (0.1 * torch.randn(D, generator=g)).to(dev())
    w1 = (torch.randn(4 * D, D, generator=g) / D ** 0.5).to(dev())
    b1 = (0.1 * torch.randn(4 * D, generator=g)).to(dev())
    w2 = (torch.randn(D, 4 * D, generator=g) / (4 * D) ** 0.5).to(dev())
    y = torch.empty(M, D, device=dev())
    z16, zpair = torch.zeros(M, D, dtype=torch.float16, device=dev()), torch.zeros(M, 2 * D, dtype=torch.uint8, device=dev())
    ops.layernorm(x, gamma, beta, y=y, out_hi=z16, out_lo=zpair, f16f8=True)
    torch.cuda.synchronize()
    y_ref = F.layer_norm(x.double(), (D,), gamma.double(), beta.double(), 1e-5)
    assert (y.double() - y_ref).abs().max().item() < 1e-5

    def decode(p16, pair):        # value represented by the planes
        k = p16.shape[1]
        pr = pair.view(p16.shape[0], k // 64, 2, 64)
        lo = pr[:, :, 0].reshape(p16.shape[0], k).view(torch.float8_e4m3fn).double() / 8192.0
        x8 = pr[:, :, 1].reshape(p16.shape[0], k).view(torch.float8_e4m3fn).double()
        return p16.double() + lo, x8
    zv, z8 = decode(z16, zpair)
    assert ((zv - y_ref).abs() <= 2.0 ** -15 * y_ref.abs() + 1e-6).all()           # fp16 + e4m3 remainder: 2^-12 * 2^-4 relative, + fp32 LN rounding
    assert ((z8 - y_ref).abs() <= 0.0625 * y_ref.abs() + 2e-3).all()                  # e4m3: 3 mantissa bits
    h16, hpair = torch.zeros(M, 4 * D, dtype=torch.float16, device=dev()), torch.zeros(M, 8 * D, dtype=torch.uint8, device=dev())
    w1p, w2p = ops.pack_f16f8(w1), ops.pack_f16f8(w2)
    ops.gemm_tc(a_hi=z16, a_lo=zpair, a_dims=(1, 1, M, D), b_hi=w1p[0], b_lo=w1p[1], k=D, n_cols=4 * D, out_w=M, ldc=4 * D, bias=b1,
                out_hi=h16, out_lo=hpair, flags=ops.GF_GELU | ops.GF_OUT_F16F8, bn=128, npass=2, lo_scale=w1p[2])
    out = torch.empty(M, D, device=dev())
    ops.gemm_tc(a_hi=h16, a_lo=hpair, a_dims=(1, 1, M, 4 * D), b_hi=w2p[0], b_lo=w2p[1], k=4 * D, n_cols=D, out_w=M, ldc=D, residual=x,
                out_f32=out, bn=128, npass=2, lo_scale=w2p[2])
    torch.cuda.synchronize()
    h_ref = F.gelu(y_ref @ w1.double().t() + b1.double())
    hv, _ = decode(h16, hpair)
    assert ((hv - h_ref).abs() <= 2.0 ** -14 * h_ref.abs() + 1e-4).all()
    ref = h_ref @ w2.double().t() + x.double()
    assert (out.double() - ref).abs().max().item() < 3e-4


@pytest.mark.parametrize("M,N,K", [(256, 256, 64), (1000, 224, 320), (3584, 768, 256), (700, 1024, 1024), (130, 32, 4096)])
def test_linear_pair_f16f8(M, N, K):
    """gemm_pair.cu (2-CTA 256x256 tiles, one accumulator, pre-scaled operands) vs fp64, with ragged M / N tails, bias and residual."""
    g = torch.Generator().manual_seed(M + N + K)
    a = torch.randn(M, K, generator=g).to(dev())
    w = (torch.randn(N, K, generator=g) / K ** 0.5).to(dev())
    bias = torch.randn(N, generator=g).to(dev())
    res = torch.randn(M, N, generator=g).to(dev())
    a16, apair = ops.pack_act_f16f8_scaled(a)
    w16, wpair, sc = ops.pack_linear_f16f8(w)
    out = torch.full((M, N), float("nan"), device=dev())
    ops.linear_f16f8(a16, apair, w16, wpair, sc, M, N, K, bias=bias, residual=res, out_f32=out)
    torch.cuda.synchronize()
    ref = a.double() @ w.double().t() + bias.double() + res.double()
    assert torch.isfinite(out).all()
    err = (out.double() - ref).abs().max().item()
    assert err < 6e-5 * K ** 0.5, f"max err {err}"


def test_linear_pair_f16f8_chain_and_planes():
    """scaled LayerNorm planes -> pair GEMM + GELU -> scaled f16f8 planes -> pair GEMM (+ residual), and the bf16 hi / lo output planes."""
    M, D = 600, 256
    g = torch.Generator().manual_seed(5)
    x = (torch.randn(M, D, generator=g) * 2 + 0.5).to(dev())
    gamma, beta = (1 + 0.1 * torch.randn(D, generator=g)).to(dev()), (0.1 * torch.randn(D, generator=g)).to(dev())
    w1 = (torch.randn(4 * D, D, generator=g) / D ** 0.5).to(dev())
    b1 = (0.1 * torch.randn(4 * D, generator=g)).to(dev())
    w2 = (torch.randn(D, 4 * D, generator=g) / (4 * D) ** 0.5).to(dev())
    z16, zpair = torch.zeros(M, D, dtype=torch.float16, device=dev()), torch.zeros(M, 2 * D, dtype=torch.uint8, device=dev())
    ops.layernorm(x, gamma, beta, out_hi=z16, out_lo=zpair, f16f8=True, scaled=True)
    y_ref = F.layer_norm(x.double(), (D,), gamma.double(), beta.double(), 1e-5)

    def decode(p16, pair):
        k = p16.shape[1]
        pr = pair.view(p16.shape[0], k // 64, 2, 64)
        lo = pr[:, :, 0].reshape(p16.shape[0], k).view(torch.float8_e4m3fn).double() / 8192.0
        return p16.double() / 64.0 + lo
    torch.cuda.synchronize()
    assert ((decode(z16, zpair) - y_ref).abs() <= 2.0 ** -15 * y_ref.abs() + 1e-6).all()
    w1p, w2p = ops.pack_linear_f16f8(w1), ops.pack_linear_f16f8(w2)
    h16, hpair = torch.zeros(M, 4 * D, dtype=torch.float16, device=dev()), torch.zeros(M, 8 * D, dtype=torch.uint8, device=dev())
    hhi, hlo = torch.zeros(M, 4 * D, dtype=torch.bfloat16, device=dev()), torch.zeros(M, 4 * D, dtype=torch.bfloat16, device=dev())
    ops.linear_f16f8(z16, zpair, w1p[0], w1p[1], w1p[2], M, 4 * D, D, bias=b1, gelu=True, out_f16=h16, out_pair=hpair, out_hi=hhi, out_lo=hlo)
    out = torch.empty(M, D, device=dev())
    ops.linear_f16f8(h16, hpair, w2p[0], w2p[1], w2p[2], M, D, 4 * D, residual=x, out_f32=out)
    torch.cuda.synchronize()
    h_ref = F.gelu(y_ref @ w1.double().t() + b1.double())
    assert ((decode(h16, hpair) - h_ref).abs() <= 2.0 ** -14 * h_ref.abs() + 1e-4).all()
    assert ((hhi.double() + hlo.double() - h_ref).abs() <= 2.0 ** -14 * h_ref.abs() + 1e-4).all()
    ref = h_ref @ w2.double().t() + x.double()
    assert (out.double() - ref).abs().max().item() < 3e-4


@pytest.mark.parametrize("N_,H,W,Cout", [(2, 64, 64, 128), (3, 37, 70, 64), (1, 256, 256, 128)])
def test_conv_in3_direct(N_, H, W, Cout):
    """Direct fp32 conv_in (NCHW RGB -> NHWC, bias, fused GroupNorm statistics) vs torch conv2d."""
    g = torch.Generator().manual_seed(H + W + Cout)
    x = torch.randn(N_, 3, H, W, generator=g).to(dev())
    w = (torch.randn(Cout, 3, 3, 3, generator=g) / 27 ** 0.5).to(dev())
    b = torch.randn(Cout, generator=g).to(dev())
    out = torch.full((N_, H, W, Cout), float("nan"), device=dev())
    sums = torch.full((N_ * 64,), float("nan"), dtype=torch.float64, device=dev())
    ops.conv_in3(x, w, b, out, gn_sums=sums)
    torch.cuda.synchronize()
    ref = F.conv2d(x.double(), w.double(), b.double(), padding=1)
    assert torch.isfinite(out).all()
    err = (out.double().permute(0, 3, 1, 2) - ref).abs().max().item()
    assert err < 1e-5, f"max err {err}"
    o = out.double().permute(0, 3, 1, 2).reshape(N_, 32, -1)
    want = torch.stack([o.sum(-1), (o * o).sum(-1)], -1).reshape(-1)
    rel = ((sums - want).abs() / (1 + want.abs())).max().item()
    assert rel < 1e-5, f"gn sums rel err {rel}"


@pytest.mark.parametrize("N_,H,W,C", [(2, 32, 32, 128), (3, 21, 50, 64), (1, 128, 128, 128)])
def test_conv_out3_direct(N_, H, W, C):
    """norm_out + swish + conv_out (C -> 3) in one direct fp32 kernel, NHWC -> NCHW, vs torch."""
    g = torch.Generator().manual_seed(H + W + C)
    x = (torch.randn(N_, C, H, W, generator=g) * 1.3 + 0.2).to(dev())
    w = (torch.randn(3, C, 3, 3, generator=g) / (9 * C) ** 0.5).to(dev())
    b = torch.randn(3, generator=g).to(dev())
    gamma, beta = (1 + 0.2 * torch.randn(C, generator=g)).to(dev()), (0.2 * torch.randn(C, generator=g)).to(dev())
    x_nhwc = x.permute(0, 2, 3, 1).contiguous()
    sums = torch.empty(N_ * 64, dtype=torch.float64, device=dev())
    mr = torch.empty(N_ * 64, dtype=torch.float32, device=dev())
    ops.groupnorm_stats(x_nhwc, sums, mr, 1e-6)
    affine = torch.empty(N_, C, 2, device=dev())
    ops.groupnorm_affine(sums, gamma, beta, affine, N_, H * W, C, 1e-6)
    out = torch.full((N_, 3, H, W), float("nan"), device=dev())
    ops.conv_out3(x_nhwc, w, b, out, affine=affine, swish=True)
    torch.cuda.synchronize()
    a = F.group_norm(x.double(), 32, gamma.double(), beta.double(), eps=1e-6)
    ref = F.conv2d(a * torch.sigmoid(a), w.double(), b.double(), padding=1)
    assert torch.isfinite(out).all()
    err = (out.double() - ref).abs().max().item()
    assert err < 2e-5, f"max err {err}"
    # no normalisation / activation
    ops.conv_out3(x_nhwc, w, b, out)
    torch.cuda.synchronize()
    assert (out.double() - F.conv2d(x.double(), w.double(), b.double(), padding=1)).abs().max().item() < 2e-5


@pytest.mark.parametrize("rows,cin", [(24, 64), (130, 320), (384, 1024)])
def test_pack_kernels_match_the_torch_statement(rows, cin):
    """bevgen_pack_split_bf16 / bevgen_pack_f16f8 / bevgen_absmax (model-load packing) are bit-identical to the torch statements of the
    formats (ops.*_torch), including saturation of the e4m3 residual plane and a zero matrix."""
    g = torch.Generator().manual_seed(rows * 7 + cin)
    for scale in (0.02, 3.0, 0.0):
        w = (torch.randn(rows, cin, generator=g) * scale).cuda()
        if scale == 3.0:
            w[0, 0] = 37.5            # drives the power-of-two weight scale; residuals of small entries then saturate nowhere, large ones may
        hi, lo = ops.split_planes(w)
        hi_r, lo_r = ops.split_planes_torch(w)
        assert torch.equal(hi.view(torch.int16), hi_r.view(torch.int16)) and torch.equal(lo.view(torch.int16), lo_r.view(torch.int16))
        assert ops.split_planes(w, npass=1)[1] is None
        for fn, ref in ((ops.pack_f16f8, ops.pack_f16f8_torch), (ops.pack_linear_f16f8, ops.pack_linear_f16f8_torch), (ops.pack_f16f8_block, ops.pack_f16f8_block_torch)):
            a16, apair, asc = fn(w)
            b16, bpair, bsc = ref(w)
            assert asc == bsc, (fn.__name__, asc, bsc)
            assert torch.equal(a16.view(torch.int16), b16.view(torch.int16)), fn.__name__
            assert torch.equal(apair, bpair), fn.__name__
