"""Thin torch-tensor -> C-ABI adapters.  torch is used for device memory and the current CUDA stream only;
all arithmetic happens inside libbevgen_b200.so."""
import ctypes as C
import math

import torch

from . import _lib
from ._lib import (GF_B_MN, GF_CAUSAL_KLIMIT, GF_CAUSAL_SKIP, GF_GELU, GF_OUT_F16F8, GF_OUT_NCHW, GF_OUT_T, PREP_IDENT, PREP_S2D, PREP_UP2,  # noqa: F401
                   EmbedArgs, GemmArgs)


class Stats:
    """Launch accounting for bench.py: number of kernel launches and algorithmic FLOPs of the GEMM launches.
    `timer`, when set, is called as timer(kind, launch_fn, flops) and must invoke launch_fn() itself (CUDA-event timing)."""
    launches = 0
    gemm_launches = 0
    gemm_flops = 0.0
    timer = None

    @classmethod
    def reset(cls):
        cls.launches, cls.gemm_launches, cls.gemm_flops = 0, 0, 0.0


def _ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _chk_cuda(*ts):
    for t in ts:
        if t is not None:
            if not t.is_cuda:
                raise RuntimeError("bevgen_b200 ops need CUDA tensors (there is no CPU path)")
            if not t.is_contiguous():
                raise RuntimeError("bevgen_b200 ops need contiguous tensors")


def split_planes(x: torch.Tensor, npass: int = 3):
    """fp32 tensor -> (hi, lo) bf16 planes with x ~= hi + lo (packing-time helper for weights): one bevgen_pack_split_bf16 launch."""
    if not x.is_cuda:
        return split_planes_torch(x, npass)
    x = x.detach().float().contiguous()
    hi = torch.empty(x.shape, dtype=torch.bfloat16, device=x.device)
    lo = torch.empty(x.shape, dtype=torch.bfloat16, device=x.device) if npass == 3 else None
    Stats.launches += 1
    _lib.check(_lib.init().bevgen_pack_split_bf16(_ptr(x), x.numel(), _ptr(hi), _ptr(lo), _stream()), "pack_split_bf16")
    return hi, lo


def split_planes_torch(x: torch.Tensor, npass: int = 3):
    """The same in torch ops (CPU tensors; the reference statement the pack kernels are tested against)."""
    hi = x.to(torch.bfloat16)
    lo = (x - hi.float()).to(torch.bfloat16) if npass == 3 else None
    return hi.contiguous(), (None if lo is None else lo.contiguous())


def _weight_scale_exp(w: torch.Tensor) -> int:
    """e with S = 2^e the largest power of two keeping max|w| * S <= 64 (bevgen_absmax: one launch + a 4-byte read-back at load time)."""
    out = torch.zeros(1, dtype=torch.float32, device=w.device)
    Stats.launches += 1
    _lib.check(_lib.init().bevgen_absmax(_ptr(w), w.numel(), _ptr(out), _stream()), "absmax")
    amax = float(out.item())
    return 6 if amax == 0.0 else min(max(6 - math.ceil(math.log2(amax)), -16), 24)


def _pack_f16f8_kernel(w2d: torch.Tensor, chunk: int, scaled16: bool):
    rows, cin = w2d.shape
    assert cin % chunk == 0
    w = w2d.detach().float().contiguous()
    s = 2.0 ** _weight_scale_exp(w)
    w16 = torch.empty((rows, cin), dtype=torch.float16, device=w.device)
    pair = torch.empty((rows, 2 * cin), dtype=torch.uint8, device=w.device)
    Stats.launches += 1
    _lib.check(_lib.init().bevgen_pack_f16f8(_ptr(w), rows, cin, chunk, s, s * 128.0 if scaled16 else 1.0, _ptr(w16), _ptr(pair), _stream()), "pack_f16f8")
    return w16, pair, 1.0 / (2.0 ** F8_ACT_LO_SHIFT * s)


F8_ACT_LO_SHIFT = 13        # conv_fused2.cu producer: lo8 = e4m3((x - fp16(x)) * 2^13), x8 = e4m3(x)
F8_ACT_SHIFT = 0


def pack_f16f8(w2d: torch.Tensor):
    """Kernel path (bevgen_pack_f16f8) of pack_f16f8_torch."""
    return _pack_f16f8_kernel(w2d, 64, False) if w2d.is_cuda else pack_f16f8_torch(w2d)


def pack_f16f8_torch(w2d: torch.Tensor):
    """fp32 weight rows [rows][cin] (cin % 64 == 0) -> (w16 [rows][cin] fp16, pair [rows][2*cin] uint8, lo_scale) for
    bevgen_conv3x3_fused_f16f8: per 64-channel chunk 64 bytes e4m3(w * S) then 64 bytes e4m3((w - fp16(w)) * S * 2^13), with
    S the largest power of two keeping max|w| * S <= 64 (so the residual plane stays <= 256; e4m3 saturates at 448)."""
    rows, cin = w2d.shape
    assert cin % 64 == 0
    w = w2d.float()
    w16 = w.to(torch.float16)
    amax = float(w.abs().max().item())
    e = 6 if amax == 0.0 else min(max(6 - math.ceil(math.log2(amax)), -16), 24)
    s = 2.0 ** e
    k = F8_ACT_LO_SHIFT - F8_ACT_SHIFT
    w8 = (w * s).clamp(-448.0, 448.0).to(torch.float8_e4m3fn)
    wlo8 = ((w - w16.float()) * (s * 2.0 ** k)).clamp(-448.0, 448.0).to(torch.float8_e4m3fn)
    pair = torch.stack([w8.view(torch.uint8).view(rows, cin // 64, 64), wlo8.view(torch.uint8).view(rows, cin // 64, 64)], 2)
    return w16.contiguous(), pair.reshape(rows, 2 * cin).contiguous(), 1.0 / (2.0 ** F8_ACT_LO_SHIFT * s)


def pack_act_f16f8(a: torch.Tensor):
    """fp32 activations [rows][k] (k % 64 == 0) -> the A-operand planes of an npass = 2 GEMM, as bevgen_layernorm_f16f8 / the
    BEVGEN_GF_OUT_F16F8 epilogue write them: (fp16 [rows][k], uint8 [rows][2k]: per 64-chunk 64 B e4m3((a - a16) * 2^13) then 64 B e4m3(a))."""
    rows, k = a.shape
    assert k % 64 == 0
    a = a.float()
    a16 = a.to(torch.float16)
    lo8 = ((a - a16.float()) * 2.0 ** F8_ACT_LO_SHIFT).clamp(-448.0, 448.0).to(torch.float8_e4m3fn)
    x8 = a.clamp(-448.0, 448.0).to(torch.float8_e4m3fn)
    pair = torch.stack([lo8.view(torch.uint8).view(rows, k // 64, 64), x8.view(torch.uint8).view(rows, k // 64, 64)], 2)
    return a16.contiguous(), pair.reshape(rows, 2 * k).contiguous()


def pack_f16f8_block(w2d: torch.Tensor):
    """Kernel path (bevgen_pack_f16f8, 32-element chunks, scaled fp16 plane) of pack_f16f8_block_torch."""
    return _pack_f16f8_kernel(w2d, 32, True) if w2d.is_cuda else pack_f16f8_block_torch(w2d)


def pack_f16f8_block_torch(w2d: torch.Tensor):
    """Weights for the 16x16-block f16f8 kernel (bevgen_conv3x3_fused_f16f8, block16=1): (w16s fp16(w * S * 2^7), pair [rows][2*cin] uint8 with,
    per 32-channel slice, 32 bytes e4m3(w * S) then 32 bytes e4m3((w - w16) * S * 2^13), lo_scale = 1 / (2^13 * S))."""
    rows, cin = w2d.shape
    assert cin % 32 == 0
    w = w2d.float()
    amax = float(w.abs().max().item())
    e = 6 if amax == 0.0 else min(max(6 - math.ceil(math.log2(amax)), -16), 24)
    s = 2.0 ** e
    w16s = (w * (s * 128.0)).to(torch.float16)
    w16 = w16s.float() / (s * 128.0)
    w8 = (w * s).clamp(-448.0, 448.0).to(torch.float8_e4m3fn)
    wlo8 = ((w - w16) * (s * 2.0 ** F8_ACT_LO_SHIFT)).clamp(-448.0, 448.0).to(torch.float8_e4m3fn)
    pair = torch.stack([w8.view(torch.uint8).view(rows, cin // 32, 32), wlo8.view(torch.uint8).view(rows, cin // 32, 32)], 2)
    return w16s.contiguous(), pair.reshape(rows, 2 * cin).contiguous(), 1.0 / (2.0 ** F8_ACT_LO_SHIFT * s)


TAPS_3X3 = [(kw - 1, kh - 1) for kh in range(3) for kw in range(3)]


def gemm_tc(*, a_hi, a_lo, a_dims, b_hi, b_lo, k, n_cols, taps=((0, 0, 0),), a_n_mul=1, a_n_zstride=0, a_c_off=0, a_c_zstride=0,
            b_k_off=0, b_k_zstride=0, b_row_zstride=0, b_row_tapstride=0, z_inner=1, z_outer=1, tile=(128, 1),
            out_w, out_h=1, out_zo_stride=0, out_zi_stride=0, ldc, bias=None, residual=None, out_f32=None, out_hi=None,
            out_lo=None, flags=0, causal_ncond=0, bn=128, npass=3, algo_flops=None, fin=None, lo_scale=0.0):
    lib = _lib.init()
    _chk_cuda(a_hi, a_lo, b_hi, b_lo, bias, residual, out_f32, out_hi, out_lo)
    g = GemmArgs()
    g.a_hi, g.a_lo = _ptr(a_hi), _ptr(a_lo)
    g.a_n, g.a_h, g.a_w, g.a_c = a_dims
    g.b_hi, g.b_lo = _ptr(b_hi), _ptr(b_lo)
    g.b_rows, g.b_cols = b_hi.shape[0], b_hi.shape[1]
    g.ntaps = len(taps)
    for i, t in enumerate(taps):
        g.tap_dx[i], g.tap_dy[i] = t[0], t[1]
        g.tap_dn[i] = t[2] if len(t) > 2 else 0
    g.a_n_mul, g.a_n_zstride, g.k = a_n_mul, a_n_zstride, k
    g.a_c_off, g.a_c_zstride, g.b_k_off, g.b_k_zstride = a_c_off, a_c_zstride, b_k_off, b_k_zstride
    g.b_row_zstride, g.b_row_tapstride = b_row_zstride, b_row_tapstride
    g.z_inner, g.z_outer = z_inner, z_outer
    g.tile_w, g.tile_h = tile
    g.out_w, g.out_h, g.n_cols = out_w, out_h, n_cols
    g.out_zo_stride, g.out_zi_stride, g.ldc = out_zo_stride, out_zi_stride, ldc
    g.bias, g.residual = _ptr(bias), _ptr(residual)
    g.out_f32, g.out_hi, g.out_lo = _ptr(out_f32), _ptr(out_hi), _ptr(out_lo)
    g.flags, g.causal_ncond, g.bn, g.npass = flags, causal_ncond, bn, npass
    g.lo_scale = float(lo_scale)
    if fin is not None:        # fused split-K finalize (decode): dict(mode, rows, counters, hi, lo, bias, gelu, resid, x, y, gamma, beta, eps)
        g.fin_mode, g.fin_gelu, g.fin_rows = fin["mode"], int(fin.get("gelu", 0)), fin["rows"]
        g.fin_bias, g.fin_resid = _ptr(fin.get("bias")), _ptr(fin.get("resid"))
        g.fin_x, g.fin_y, g.fin_hi, g.fin_lo = _ptr(fin.get("x")), _ptr(fin.get("y")), _ptr(fin["hi"]), _ptr(fin.get("lo"))
        g.fin_gamma, g.fin_beta, g.fin_eps = _ptr(fin.get("gamma")), _ptr(fin.get("beta")), float(fin.get("eps", 1e-5))
        g.fin_counters = _ptr(fin["counters"])
    flops = algo_flops if algo_flops is not None else 2.0 * z_inner * z_outer * out_w * out_h * n_cols * k * len(taps)
    Stats.launches += 1
    Stats.gemm_launches += 1
    Stats.gemm_flops += flops
    if Stats.timer is not None:
        Stats.timer("gemm_tc", lambda: _lib.check(lib.bevgen_gemm_tc(C.byref(g), _stream()), "bevgen_gemm_tc"), flops)
    else:
        _lib.check(lib.bevgen_gemm_tc(C.byref(g), _stream()), "bevgen_gemm_tc")


def groupnorm_stats(x_nhwc, ws_sums, mean_rstd, eps=1e-6):
    lib = _lib.init()
    Stats.launches += 2
    _chk_cuda(x_nhwc, ws_sums, mean_rstd)
    n, c = x_nhwc.shape[0], x_nhwc.shape[-1]
    pixels = x_nhwc.numel() // (n * c)
    _lib.check(lib.bevgen_groupnorm_stats(_ptr(x_nhwc), n, pixels, c, eps, _ptr(ws_sums), _ptr(mean_rstd), _stream()), "groupnorm_stats")


def prep_operand(x_nhwc, out_hi, out_lo, mean_rstd=None, gamma=None, beta=None, swish=False, mode=PREP_IDENT):
    lib = _lib.init()
    Stats.launches += 1
    _chk_cuda(x_nhwc, out_hi, out_lo, mean_rstd, gamma, beta)
    n, h, w, c = x_nhwc.shape
    _lib.check(lib.bevgen_prep_operand(_ptr(x_nhwc), n, h, w, c, _ptr(mean_rstd), _ptr(gamma), _ptr(beta), int(swish), mode,
                                       _ptr(out_hi), _ptr(out_lo), _stream()), "prep_operand")


def im2col3x3(x_nchw, out_hi, out_lo):
    lib = _lib.init()
    Stats.launches += 1
    _chk_cuda(x_nchw, out_hi, out_lo)
    n, cin, h, w = x_nchw.shape
    _lib.check(lib.bevgen_im2col3x3(_ptr(x_nchw), n, cin, h, w, _ptr(out_hi), _ptr(out_lo), _stream()), "im2col3x3")


def transpose_f32(src, dst, n, r, c):
    lib = _lib.init()
    Stats.launches += 1
    _chk_cuda(src, dst)
    _lib.check(lib.bevgen_transpose_f32(_ptr(src), _ptr(dst), n, r, c, _stream()), "transpose_f32")


def softmax_rows(s, out_hi, out_lo, scale, out_ld=None):
    lib = _lib.init()
    Stats.launches += 1
    _chk_cuda(s, out_hi, out_lo)
    cols = s.shape[-1]
    _lib.check(lib.bevgen_softmax_rows(_ptr(s), s.numel() // cols, cols, float(scale), _ptr(out_hi), _ptr(out_lo),
                                       cols if out_ld is None else out_ld, _stream()), "softmax_rows")


def row_sqnorm(x, out):
    lib = _lib.init()
    Stats.launches += 1
    _chk_cuda(x, out)
    _lib.check(lib.bevgen_row_sqnorm(_ptr(x), x.shape[0], x.shape[1], _ptr(out), _stream()), "row_sqnorm")


def vq_nearest(z, codebook, code_sqnorm, ws_zz, idx, zq=None):
    lib = _lib.init()
    Stats.launches += 2
    _chk_cuda(z, codebook, code_sqnorm, ws_zz, idx, zq)
    assert idx.dtype == torch.int64
    _lib.check(lib.bevgen_vq_nearest(_ptr(z), _ptr(codebook), _ptr(code_sqnorm), z.shape[0], codebook.shape[0], codebook.shape[1],
                                     _ptr(ws_zz), _ptr(idx), _ptr(zq), _stream()), "vq_nearest")


def codebook_gather(codebook, idx, out):
    lib = _lib.init()
    Stats.launches += 1
    _chk_cuda(codebook, idx, out)
    assert idx.dtype == torch.int64
    _lib.check(lib.bevgen_codebook_gather(_ptr(codebook), _ptr(idx), idx.numel(), codebook.shape[1], codebook.shape[0], _ptr(out),
                                          _stream()), "codebook_gather")


def conv_in3(x_nchw, weight_oihw, bias, out_nhwc, gn_sums=None):
    """Encoder.conv_in for 3-channel inputs: direct fp32 conv NCHW -> NHWC (+ GroupNorm statistics of the output)."""
    lib = _lib.init()
    _chk_cuda(x_nchw, weight_oihw, bias, out_nhwc, gn_sums)
    n, cin, h, w = x_nchw.shape
    cout = weight_oihw.shape[0]
    assert cin == 3 and tuple(weight_oihw.shape) == (cout, 3, 3, 3)
    Stats.launches += 1
    Stats.gemm_launches += 1
    flops = 2.0 * n * h * w * cout * 27
    Stats.gemm_flops += flops
    call = lambda: _lib.check(lib.bevgen_conv_in3(_ptr(x_nchw), _ptr(weight_oihw), _ptr(bias), _ptr(out_nhwc), _ptr(gn_sums), n, h, w, cout, _stream()),
                              "conv_in3")
    if Stats.timer is not None:
        Stats.timer("conv_small", call, flops)
    else:
        call()


def conv_out3(x_nhwc, weight_oihw, bias, out_nchw, affine=None, swish=False):
    """Decoder.conv_out for 3-channel outputs: GroupNorm-apply + swish + direct fp32 3x3 conv, NHWC -> NCHW."""
    lib = _lib.init()
    _chk_cuda(x_nhwc, weight_oihw, bias, out_nchw, affine)
    n, h, w, c = x_nhwc.shape
    assert tuple(weight_oihw.shape) == (3, c, 3, 3)
    Stats.launches += 1
    Stats.gemm_launches += 1
    flops = 2.0 * n * h * w * 3 * c * 9
    Stats.gemm_flops += flops
    call = lambda: _lib.check(lib.bevgen_conv_out3(_ptr(x_nhwc), n, h, w, c, _ptr(affine), int(swish), _ptr(weight_oihw), _ptr(bias), _ptr(out_nchw),
                                                   _stream()), "conv_out3")
    if Stats.timer is not None:
        Stats.timer("conv_small", call, flops)
    else:
        call()


def to_uint8_hwc(x_nchw, out=None):
    """fp32 (N, C, H, W) in [0, 1] on the GPU -> uint8 (N, H, W, C)."""
    lib = _lib.init()
    Stats.launches += 1
    n, c, h, w = x_nchw.shape
    out = torch.empty((n, h, w, c), dtype=torch.uint8, device=x_nchw.device) if out is None else out
    _chk_cuda(x_nchw, out)
    _lib.check(lib.bevgen_to_uint8_hwc(_ptr(x_nchw), _ptr(out), n, c, h * w, _stream()), "to_uint8_hwc")
    return out


def denormalize(x_nchw, out, mean, std):
    lib = _lib.init()
    Stats.launches += 1
    _chk_cuda(x_nchw, out)
    n, c, h, w = x_nchw.shape
    m = (C.c_float * 3)(*mean)
    s = (C.c_float * 3)(*std)
    _lib.check(lib.bevgen_denormalize(_ptr(x_nchw), _ptr(out), n, c, h * w, m, s, _stream()), "denormalize")


def layernorm(x, gamma, beta, y=None, out_hi=None, out_lo=None, eps=1e-5, rows=None, row_stride=None, f16f8=False, scaled=False):
    """f16f8: out_hi / out_lo receive the fp16 plane and the e4m3 pair plane (operands of an npass = 2 GEMM); scaled: the fp16 plane holds
    2^6 * y (operand convention of linear_f16f8)."""
    lib = _lib.init()
    Stats.launches += 1
    _chk_cuda(gamma, beta, y, out_hi, out_lo)
    d = gamma.numel()
    rows = x.numel() // d if rows is None else rows
    row_stride = d if row_stride is None else row_stride
    if f16f8:
        _lib.check(lib.bevgen_layernorm_f16f8(_ptr(x), rows, d, row_stride, _ptr(gamma), _ptr(beta), eps, _ptr(y), _ptr(out_hi), _ptr(out_lo),
                                              1 if scaled else 0, _stream()), "layernorm_f16f8")
    else:
        _lib.check(lib.bevgen_layernorm(_ptr(x), rows, d, row_stride, _ptr(gamma), _ptr(beta), eps, _ptr(y), _ptr(out_hi), _ptr(out_lo), _stream()),
                   "layernorm")


def pack_linear_f16f8(w2d: torch.Tensor):
    """Kernel path (bevgen_pack_f16f8, scaled fp16 plane) of pack_linear_f16f8_torch."""
    return _pack_f16f8_kernel(w2d, 64, True) if w2d.is_cuda else pack_linear_f16f8_torch(w2d)


def pack_linear_f16f8_torch(w2d: torch.Tensor):
    """nn.Linear weight [out][in] (in % 64 == 0) -> operands of linear_f16f8: (w16s = fp16(w * S * 2^7), pair [out][2*in] uint8 with, per
    64-element chunk, 64 bytes e4m3(w * S) then 64 bytes e4m3((w - w16) * S * 2^13), out_scale = 1 / (2^13 * S))."""
    rows, cin = w2d.shape
    assert cin % 64 == 0
    w = w2d.float()
    amax = float(w.abs().max().item())
    e = 6 if amax == 0.0 else min(max(6 - math.ceil(math.log2(amax)), -16), 24)
    s = 2.0 ** e
    w16s = (w * (s * 128.0)).to(torch.float16)
    w16 = w16s.float() / (s * 128.0)
    w8 = (w * s).clamp(-448.0, 448.0).to(torch.float8_e4m3fn)
    wlo8 = ((w - w16) * (s * 2.0 ** F8_ACT_LO_SHIFT)).clamp(-448.0, 448.0).to(torch.float8_e4m3fn)
    pair = torch.stack([w8.view(torch.uint8).view(rows, cin // 64, 64), wlo8.view(torch.uint8).view(rows, cin // 64, 64)], 2)
    return w16s.contiguous(), pair.reshape(rows, 2 * cin).contiguous(), 1.0 / (2.0 ** F8_ACT_LO_SHIFT * s)


def pack_act_f16f8_scaled(a: torch.Tensor):
    """fp32 activations [rows][k] -> the scaled A planes of linear_f16f8 (fp16(a * 2^6), pair plane as pack_act_f16f8 against the unscaled a16)."""
    rows, k = a.shape
    assert k % 64 == 0
    a = a.float()
    a16s = (a * 64.0).to(torch.float16)
    lo8 = ((a - a16s.float() / 64.0) * 2.0 ** F8_ACT_LO_SHIFT).clamp(-448.0, 448.0).to(torch.float8_e4m3fn)
    x8 = a.clamp(-448.0, 448.0).to(torch.float8_e4m3fn)
    pair = torch.stack([lo8.view(torch.uint8).view(rows, k // 64, 64), x8.view(torch.uint8).view(rows, k // 64, 64)], 2)
    return a16s.contiguous(), pair.reshape(rows, 2 * k).contiguous()


def linear_f16f8(a16, apair, w16, wpair, out_scale, M, N, K, bias=None, gelu=False, residual=None, out_f32=None, out_hi=None, out_lo=None,
                 out_f16=None, out_pair=None):
    """y = act(a @ w.T + bias) (+ residual) on the 2-CTA f16f8 GEMM (gemm_pair.cu): operands from layernorm(..., f16f8=True, scaled=True) /
    a previous linear_f16f8(out_f16=, out_pair=) and pack_linear_f16f8.  Outputs: fp32 and/or bf16 hi/lo planes and/or scaled f16f8 planes."""
    lib = _lib.init()
    Stats.launches += 1
    _chk_cuda(a16, apair, w16, wpair, bias, residual, out_f32, out_hi, out_lo, out_f16, out_pair)
    flops = 2.0 * M * N * K
    Stats.gemm_launches += 1
    Stats.gemm_flops += flops
    call = lambda: _lib.check(lib.bevgen_linear_f16f8(_ptr(a16), _ptr(apair), _ptr(w16), _ptr(wpair), M, N, K, out_scale, _ptr(bias),
                                                      1 if gelu else 0, _ptr(residual), _ptr(out_f32), _ptr(out_hi), _ptr(out_lo), _ptr(out_f16),
                                                      _ptr(out_pair), _stream()), "linear_f16f8")
    if Stats.timer is not None:
        Stats.timer("linear_f16f8", call, flops)
    else:
        call()


def embed_assemble(args: "EmbedArgs"):
    lib = _lib.init()
    Stats.launches += 1
    _lib.check(lib.bevgen_embed_assemble(C.byref(args), _stream()), "embed_assemble")


def attn_softmax(S, bias, mask_u8, out_hi, out_lo, L, scale, layout=None, heads=1, block=16):
    """layout: optional uint8 [heads][nb][nb] per-head block layout (density < 1); rows of S are ordered [batch][head][query]."""
    lib = _lib.init()
    Stats.launches += 1
    _chk_cuda(S, bias, mask_u8, out_hi, out_lo, layout)
    Lk = S.shape[-1]
    _lib.check(lib.bevgen_attn_softmax(_ptr(S), _ptr(bias), _ptr(mask_u8), S.numel() // Lk, L, Lk, float(scale), _ptr(out_hi), _ptr(out_lo),
                                       _ptr(layout), heads, block, 0 if layout is None else layout.shape[-1], _stream()), "attn_softmax")


def tile_attention_bias(bias: torch.Tensor, scale: float) -> torch.Tensor:
    """[L][L] fp32 camera bias -> the tiled, pre-scaled fp16 table bevgen_attn_fused_fwd reads:
    out[qt][kt][p][u][r][e] = fp16(scale * log2(e) * bias[128 qt + r][128 kt + 32 p + 8 u + e])."""
    L = bias.shape[0]
    assert bias.shape == (L, L) and L % 128 == 0
    n = L // 128
    b = (bias.float() * (scale * 1.4426950408889634)).to(torch.float16)
    return b.view(n, 128, n, 4, 4, 8).permute(0, 2, 3, 4, 1, 5).contiguous()


def layout_to_tiles64(layout: torch.Tensor, block: int, L: int) -> torch.Tensor:
    """DeepSpeed block layout uint8 [H][nb][nb] (block in 16 / 32 / 64 / 128, nb * block >= L) -> the fused attention kernel's table
    int64 [H][L/128][L/128]: bit 8*rb + kb = sub-block (16 query rows rb, 16 keys kb) of the 128 x 128 tile is attended."""
    assert block in (16, 32, 64, 128) and L % 128 == 0
    H = layout.shape[0]
    n16, rep, nt = L // 16, block // 16, L // 128
    fine = (layout != 0).repeat_interleave(rep, 1).repeat_interleave(rep, 2)[:, :n16, :n16]        # [H][L/16][L/16]
    t = fine.reshape(H, nt, 8, nt, 8).permute(0, 1, 3, 2, 4).reshape(H, nt, nt, 64).to(torch.int64)
    return (t << torch.arange(64, device=layout.device, dtype=torch.int64)).sum(-1).contiguous()      # bit 63 wraps into the sign


def attn_fused_fwd(qkv_hi, qkv_lo, B, L, H, d, n_cond, bias_f16, y, x1, scale, npass, algo_flops=0.0, layout64=None, out_hi=None, out_lo=None):
    lib = _lib.init()
    _chk_cuda(qkv_hi, qkv_lo, bias_f16, y, x1, layout64, out_hi, out_lo)
    Stats.launches += 1
    call = lambda: _lib.check(lib.bevgen_attn_fused_fwd(_ptr(qkv_hi), _ptr(qkv_lo), B, L, H, d, n_cond, _ptr(bias_f16), _ptr(y), _ptr(x1),
                                                        float(scale), npass, _ptr(layout64), _ptr(out_hi), _ptr(out_lo), _stream()), "attn_fused_fwd")
    if Stats.timer is not None:
        Stats.timer("attn_fused", call, algo_flops)
    else:
        call()


def conv3x3_halo(a_hi, a_lo, dims, w_hi, w_lo, cout, bias, out, residual=None, gn_sums=None, npass=3):
    """3x3 s1 'same' conv on NHWC planes via the halo-tile kernel; optional fused GroupNorm statistics of the output."""
    lib = _lib.init()
    _chk_cuda(a_hi, a_lo, w_hi, w_lo, bias, out, residual, gn_sums)
    n, h, w, cin = dims
    flops = 2.0 * n * h * w * cout * cin * 9
    Stats.launches += 1
    Stats.gemm_launches += 1
    Stats.gemm_flops += flops
    call = lambda: _lib.check(lib.bevgen_conv3x3_halo(_ptr(a_hi), _ptr(a_lo), n, h, w, cin, _ptr(w_hi), _ptr(w_lo), w_hi.shape[0], cout, _ptr(bias),
                                                      _ptr(residual), _ptr(out), _ptr(gn_sums), npass, _stream()), "conv3x3_halo")
    if Stats.timer is not None:
        Stats.timer("conv_halo", call, flops)
    else:
        call()


def groupnorm_finalize(sums, mean_rstd, n, pixels, c, eps=1e-6):
    lib = _lib.init()
    Stats.launches += 1
    _chk_cuda(sums, mean_rstd)
    _lib.check(lib.bevgen_groupnorm_finalize(_ptr(sums), n, pixels, c, eps, _ptr(mean_rstd), _stream()), "groupnorm_finalize")


def conv3x3_fused(x, w_hi, w_lo, cout, bias, out, affine=None, swish=False, up2=False, residual=None, gn_sums=None, npass=3, two_cta=False,
                  block16=False):
    """3x3 s1 'same' conv straight from the fp32 NHWC activation `x` (GroupNorm-apply/swish/split/upsample fused into the operand path)."""
    lib = _lib.init()
    _chk_cuda(x, w_hi, w_lo, bias, out, affine, residual, gn_sums)
    n, h, w, cout_ = out.shape
    cin = x.shape[-1]
    flops = 2.0 * n * h * w * cout * cin * 9
    Stats.launches += 1
    Stats.gemm_launches += 1
    Stats.gemm_flops += flops
    call = lambda: _lib.check(lib.bevgen_conv3x3_fused(_ptr(x), n, h, w, cin, _ptr(affine), int(swish), int(up2), _ptr(w_hi), _ptr(w_lo), w_hi.shape[0],
                                                       cout, _ptr(bias), _ptr(residual), _ptr(out), _ptr(gn_sums),
                                                       npass | (0x200 if block16 else (0x100 if two_cta else 0)), _stream()), "conv3x3_fused")
    if Stats.timer is not None:
        Stats.timer("conv_fused", call, flops)
    else:
        call()


def conv3x3_fused_f16f8(x, w16, w8pair, lo_scale, cout, bias, out, affine=None, swish=False, up2=False, residual=None, gn_sums=None, block16=False):
    """conv3x3_fused with the fp32-equivalent product formed as one fp16 MMA + two e4m3 MMAs (weights from pack_f16f8, or from
    pack_f16f8_block when block16 selects the weight-stationary 16x16-block kernel: GroupNorm-ed inputs only, |x| < 1024)."""
    lib = _lib.init()
    _chk_cuda(x, w16, w8pair, bias, out, affine, residual, gn_sums)
    n, h, w, _ = out.shape
    cin = x.shape[-1]
    flops = 2.0 * n * h * w * cout * cin * 9
    Stats.launches += 1
    Stats.gemm_launches += 1
    Stats.gemm_flops += flops
    call = lambda: _lib.check(lib.bevgen_conv3x3_fused_f16f8(_ptr(x), n, h, w, cin, _ptr(affine), int(swish), int(up2), _ptr(w16), _ptr(w8pair),
                                                             w16.shape[0], cout, float(lo_scale), _ptr(bias), _ptr(residual), _ptr(out),
                                                             _ptr(gn_sums), int(block16), _stream()), "conv3x3_fused_f16f8")
    if Stats.timer is not None:
        Stats.timer("conv_fused", call, flops)
    else:
        call()


def groupnorm_affine(sums, gamma, beta, affine, n, pixels, c, eps=1e-6):
    lib = _lib.init()
    Stats.launches += 1
    _chk_cuda(sums, gamma, beta, affine)
    _lib.check(lib.bevgen_groupnorm_affine(_ptr(sums), _ptr(gamma), _ptr(beta), n, pixels, c, eps, _ptr(affine), _stream()), "groupnorm_affine")


def mg_head_planes(src, src_ld, src_col0, n_src, out_hi, out_lo, batch, dst_rows, heads, null_vec=None, scale=None, src_batch_rows=None,
                   dst_ld=None, dst_col0=0, dst_batch_rows=None):
    """MaskGit attention operand planes (bevgen_mg_head_planes): head split, optional null row, optional cosine-sim normalisation."""
    lib = _lib.init()
    Stats.launches += 1
    _chk_cuda(src, out_hi, out_lo, null_vec, scale)
    _lib.check(lib.bevgen_mg_head_planes(_ptr(src), src_ld, src_col0, n_src, n_src if src_batch_rows is None else src_batch_rows, _ptr(null_vec),
                                         _ptr(scale), _ptr(out_hi), _ptr(out_lo), batch, dst_rows, dst_rows if dst_batch_rows is None else dst_batch_rows,
                                         64 * heads if dst_ld is None else dst_ld, dst_col0,
                                         0 if null_vec is None else 1, heads, _stream()), "mg_head_planes")


def mg_geglu_ln(h, gamma, out_hi, out_lo, rows, f, f_pad, eps=1e-5, h_ld=None, f16f8=False):
    lib = _lib.init()
    Stats.launches += 1
    _chk_cuda(h, gamma, out_hi, out_lo)
    _lib.check(lib.bevgen_mg_geglu_ln(_ptr(h), 2 * f if h_ld is None else h_ld, _ptr(gamma), _ptr(out_hi), _ptr(out_lo), rows, f, f_pad, eps,
                                      1 if f16f8 else 0, _stream()), "mg_geglu_ln")


def ray_embed_add(h_nhwc, intrinsics_inv, extrinsics_inv, pixel, img_w, cam_w):
    """Stage-1 geometric embedding added in place to the NHWC encoder output (bevgen_ray_embed_add)."""
    lib = _lib.init()
    Stats.launches += 1
    _chk_cuda(h_nhwc, intrinsics_inv, extrinsics_inv, pixel, img_w, cam_w)
    n, hh, ww, d = h_nhwc.shape
    if intrinsics_inv.numel() != n * 9 or extrinsics_inv.numel() != n * 16 or pixel.numel() != hh * ww * 3:
        raise ValueError("ray_embed_add: camera matrices / pixel plane do not match the latent shape")
    _lib.check(lib.bevgen_ray_embed_add(_ptr(h_nhwc), _ptr(intrinsics_inv), _ptr(extrinsics_inv), _ptr(pixel), _ptr(img_w), _ptr(cam_w), n, hh * ww, d,
                                        _stream()), "ray_embed_add")


def mg_sample(logits, uniform, ids, top_k, inv_temperature, mask_id, scores=None):
    """bevgen_mg_sample: ids[row] <- argmax(top-k(logits) * inv_temperature + gumbel(uniform)) where ids[row] == mask_id (in place);
    scores (optional, fp32 [rows]) <- 1 - softmax(logits)[pred] at masked rows, -1e5 elsewhere."""
    _chk_cuda(logits, uniform, ids)
    rows, V = logits.numel() // logits.shape[-1], logits.shape[-1]
    assert logits.dtype == torch.float32 and uniform.dtype == torch.float32 and ids.dtype == torch.int64 and ids.numel() == rows and uniform.numel() == rows * V
    Stats.launches += 1
    _lib.check(_lib.init().bevgen_mg_sample(_ptr(logits), _ptr(uniform), _ptr(ids), _ptr(scores), rows, V, int(top_k), float(inv_temperature), int(mask_id),
                                           _stream()), "mg_sample")


def mg_remask(scores, ids, n_mask, mask_id, uniform=None, noise_scale=0.0, init_ids=None):
    """bevgen_mg_remask: per camera row [hw] of `ids` (in place), the n_mask largest scores (+ (uniform - 0.5) * noise_scale) -> mask_id, then
    positions where init_ids != mask_id are restored."""
    _chk_cuda(scores, ids)
    hw = ids.shape[-1]
    rows = ids.numel() // hw
    assert scores.dtype == torch.float32 and scores.numel() == rows * hw and ids.dtype == torch.int64
    Stats.launches += 1
    _lib.check(_lib.init().bevgen_mg_remask(_ptr(scores), _ptr(uniform), float(noise_scale), _ptr(ids), _ptr(init_ids), rows, hw, int(n_mask), int(mask_id),
                                           _stream()), "mg_remask")

