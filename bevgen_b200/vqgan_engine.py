"""Stage-1 VQGAN inference engine: Encoder -> quant_conv -> nearest-code VQ -> post_quant_conv -> Decoder, executed
as a flat program of C-ABI kernel launches over NHWC activations.

Reference behaviour (under /root/reference/multi_view_generation/modules/stage1): Encoder.forward model.py:406-433,
Decoder.forward :506-537, ResnetBlock :117-137, AttnBlock :168-192, Downsample :68-75, Upsample :49-53,
VQModel.encode/decode vqgan.py:84-121, VectorQuantizer2 quantize.py:271-329.

Data layout in HBM: the residual stream of every level is fp32 NHWC; each conv consumes bf16 operand planes
(hi [+ lo]) written by the `prep` kernel (GroupNorm-apply + swish + split [+ upsample / space-to-depth]) and
writes fp32 NHWC with bias and residual fused in the epilogue.  `precision="fp32x3"` (default) runs the bf16x3
split product (fp32-equivalent, meets the 1e-3 pixel / bit-exact-token bar); `precision="bf16"` is single pass.
torch is used for memory only.
"""
import math

import torch

from . import ops


def _pad_rows(w2d: torch.Tensor, rows: int) -> torch.Tensor:
    if w2d.shape[0] >= rows:
        return w2d.contiguous()
    out = torch.zeros(rows, w2d.shape[1], dtype=w2d.dtype, device=w2d.device)
    out[: w2d.shape[0]] = w2d
    return out


def _bn_for(cout: int) -> int:
    return 128 if cout >= 128 else (64 if cout >= 64 else 16)


class PackedConv:
    """Weights of one Conv2d re-laid as [tap][cout][cin] bf16 planes (K-major rows), fp32 bias."""

    def __init__(self, weight, bias, npass, device, im2col_in=False, f16f8=False):
        cout, cin, kh, kw = weight.shape
        self.cout, self.cin, self.ntaps = cout, cin, kh * kw
        self.bn = _bn_for(cout)
        w = weight.to(device=device, dtype=torch.float32)
        if im2col_in:      # conv_in: k = (kh*3+kw)*cin + c, padded to 64 (elementwise.cu im2col3x3_kernel)
            w2 = w.permute(0, 2, 3, 1).reshape(cout, kh * kw * cin)
            w2 = torch.cat([w2, torch.zeros(cout, 64 - w2.shape[1], device=device)], 1)
            self.ntaps, self.k = 1, 64
        else:
            w2 = w.permute(2, 3, 0, 1).reshape(kh * kw * cout, cin)
            self.k = cin
        n_tiles = (cout + self.bn - 1) // self.bn
        w2 = _pad_rows(w2, (self.ntaps - 1) * cout + max(n_tiles * self.bn, ((cout + 127) // 128) * 128))
        self.hi, self.lo = ops.split_planes(w2, npass)
        self.f16f8 = None         # (w16, e4m3 pair, lo_scale) for the fp16 + 2 x e4m3 conv kernel, packed on request
        if f16f8 and not im2col_in and self.ntaps == 9 and cin % 64 == 0 and cout % 32 == 0:
            self.f16f8 = ops.pack_f16f8(w2)
        self.bias = None if bias is None else bias.to(device=device, dtype=torch.float32).contiguous()


class VQGANEngine:
    def __init__(self, state_dict, ddconfig, n_embed=1024, embed_dim=256, device="cuda", precision="fp32x3"):
        # "fp32x3": every GEMM is the bf16x3 split product.  "f16f8" (parity mode, default of bench.py): the 3x3 stride-1 convs (93 % of
        # the FLOPs) form the same fp32-equivalent product as 1 fp16 + 2 e4m3 MMAs (2/3 of the tensor time); the rest stays bf16x3.
        assert precision in ("fp32x3", "f16f8", "bf16")
        self.npass = 1 if precision == "bf16" else 3
        self.f16f8 = precision == "f16f8"
        self.precision = precision
        self.dev = torch.device(device)
        self.dd = dict(ddconfig)
        self.n_embed, self.embed_dim = n_embed, embed_dim
        self.sd = state_dict
        self.use_direct_conv_out = True # RGB conv_out: norm_out + swish + direct CUDA-core conv in one kernel (conv_small.cu)
        self.use_direct_conv_in = True  # RGB conv_in: direct CUDA-core kernel (conv_small.cu) instead of im2col + tcgen05 GEMM
        self.use_block16 = True       # bf16 mode: 16x16-block weight-stationary conv kernel on the large feature maps
        self.use_two_cta = True       # cta_group::2 conv kernel (clusters of two CTAs share each weight tile)
        self.use_fused = True         # conv reads fp32 activations directly; GroupNorm/swish/split/upsample fused into its operand path
        self.use_halo = True          # halo-tile conv kernel for 3x3 stride-1 convs (Cin % 64 == 0)
        self.w = {}
        self._pack()

    # ------------------------------------------------------------------ packing
    def _conv(self, name, im2col_in=False):
        if name not in self.w:
            self.w[name] = PackedConv(self.sd[f"{name}.weight"], self.sd.get(f"{name}.bias"), self.npass, self.dev, im2col_in, self.f16f8)
        return self.w[name]

    def _norm(self, name):
        key = ("norm", name)
        if key not in self.w:
            self.w[key] = (self.sd[f"{name}.weight"].to(self.dev, torch.float32).contiguous(),
                           self.sd[f"{name}.bias"].to(self.dev, torch.float32).contiguous())
        return self.w[key]

    def _qkv(self, name):
        key = ("qkv", name)
        if key not in self.w:
            wq = torch.cat([self.sd[f"{name}.{p}.weight"] for p in ("q", "k", "v")], 0)
            bq = torch.cat([self.sd[f"{name}.{p}.bias"] for p in ("q", "k", "v")], 0)
            self.w[key] = PackedConv(wq, bq, self.npass, self.dev)
        return self.w[key]

    def _pack(self):
        sd = self.sd
        for k in list(sd.keys()):
            if k.endswith(".weight") and sd[k].dim() == 4:
                name = k[: -len(".weight")]
                if name.split(".")[-1] in ("q", "k", "v"):
                    self._qkv(name.rsplit(".", 1)[0])
                elif name == "encoder.conv_in":
                    self._conv(name, im2col_in=True)
                else:
                    self._conv(name)
            elif k.endswith(".weight") and sd[k].dim() == 1:
                self._norm(k[: -len(".weight")])
        self.codebook = sd["quantize.embedding.weight"].to(self.dev, torch.float32).contiguous()
        self.code_sqnorm = torch.empty(self.codebook.shape[0], device=self.dev)
        ops.row_sqnorm(self.codebook, self.code_sqnorm)
        self.has = lambda name: f"{name}.weight" in sd
        self.nlev = len(self.dd["ch_mult"])
        self.nres = self.dd["num_res_blocks"]

    # ------------------------------------------------------------------ primitive steps
    def _planes(self, shape):
        hi = torch.empty(shape, dtype=torch.bfloat16, device=self.dev)
        lo = torch.empty(shape, dtype=torch.bfloat16, device=self.dev) if self.npass == 3 else None
        return hi, lo

    def _prep(self, x, norm=None, swish=False, mode=ops.PREP_IDENT):
        n, h, w, c = x.shape
        if mode == ops.PREP_UP2:
            shape = (n, 2 * h, 2 * w, c)
        elif mode == ops.PREP_S2D:
            shape = (n * 4, h // 2, w // 2, c)
        else:
            shape = (n, h, w, c)
        hi, lo = self._planes(shape)
        if norm is not None:
            gamma, beta = self._norm(norm)
            mr = torch.empty(n * 64, dtype=torch.float32, device=self.dev)
            sums = getattr(x, "_gn_sums", None)
            if sums is not None:          # statistics were accumulated by the producing conv's epilogue
                ops.groupnorm_finalize(sums, mr, n, h * w, c, 1e-6)
            else:
                ws = torch.empty(n * 64, dtype=torch.float64, device=self.dev)
                ops.groupnorm_stats(x, ws, mr, 1e-6)
            ops.prep_operand(x, hi, lo, mr, gamma, beta, swish=swish, mode=mode)
        else:
            ops.prep_operand(x, hi, lo, mode=mode)
        return hi, lo

    def _gemm_conv(self, planes, pc: PackedConv, taps, geom, out_hw, residual=None, a_n_mul=1, out_planes=False, nchw=False, algo_flops=None):
        """planes: (hi, lo) of logical shape geom=(a_n,a_h,a_w,a_c); output pixels out_hw=(N,H,W)."""
        hi, lo = planes
        N, H, W = out_hw
        tw = 16 if W >= 16 else 8
        kw = dict(a_hi=hi, a_lo=lo, a_dims=geom, b_hi=pc.hi, b_lo=pc.lo, k=pc.k, n_cols=pc.cout, taps=taps, a_n_mul=a_n_mul,
                  b_row_tapstride=pc.cout, z_outer=N, tile=(tw, 128 // tw), out_w=W, out_h=H, out_zo_stride=H * W * pc.cout,
                  ldc=pc.cout, bias=pc.bias, residual=residual, bn=pc.bn, npass=self.npass, algo_flops=algo_flops)
        if out_planes:
            oh, ol = self._planes((N, H, W, pc.cout))
            ops.gemm_tc(out_hi=oh, out_lo=ol, **kw)
            return oh, ol
        if nchw:
            out = torch.empty((N, pc.cout, H, W), dtype=torch.float32, device=self.dev)
            ops.gemm_tc(out_f32=out, flags=ops.GF_OUT_NCHW, **kw)
        else:
            out = torch.empty((N, H, W, pc.cout), dtype=torch.float32, device=self.dev)
            ops.gemm_tc(out_f32=out, **kw)
        return out

    _TAPS3 = [(dx, dy, 0) for dx, dy in ops.TAPS_3X3]
    _TAPS1 = [(0, 0, 0)]
    _TAPS_S2 = [(kw // 2, kh // 2, (kh & 1) * 2 + (kw & 1)) for kh in range(3) for kw in range(3)]

    def _conv3x3_planes(self, planes, pc, dims, residual=None, nchw=False):
        """3x3 s1 'same' conv on operand planes: halo-tile kernel (+ fused GroupNorm statistics of the output) when eligible."""
        n, h, w, c = dims
        if self.use_halo and not nchw and c % 64 == 0 and pc.cout % 32 == 0 and pc.ntaps == 9:
            out = torch.empty((n, h, w, pc.cout), dtype=torch.float32, device=self.dev)
            sums = torch.empty(n * 64, dtype=torch.float64, device=self.dev) if pc.cout >= 128 else None   # fused stats need >= 4 ch / group
            ops.conv3x3_halo(planes[0], planes[1], dims, pc.hi, pc.lo, pc.cout, pc.bias, out, residual=residual, gn_sums=sums, npass=self.npass)
            if sums is not None:
                out._gn_sums = sums
            return out
        return self._gemm_conv(planes, pc, self._TAPS3, dims, (n, h, w), residual, nchw=nchw)

    def _conv3x3_fused(self, x, pc, norm, swish, residual, up2=False):
        """conv straight from the fp32 activation: GroupNorm-apply + swish + split (+ 2x upsample) happen in the conv's operand path."""
        n, h, w, c = x.shape
        if up2:
            h, w = 2 * h, 2 * w
        affine = None
        if norm is not None:
            gamma, beta = self._norm(norm)
            sums = getattr(x, "_gn_sums", None)
            if sums is None:
                sums = torch.empty(n * 64, dtype=torch.float64, device=self.dev)
                mr = torch.empty(n * 64, dtype=torch.float32, device=self.dev)
                ops.groupnorm_stats(x, sums, mr, 1e-6)
            affine = torch.empty((n, c, 2), dtype=torch.float32, device=self.dev)
            ops.groupnorm_affine(sums, gamma, beta, affine, n, x.shape[1] * x.shape[2], c, 1e-6)
        out = torch.empty((n, h, w, pc.cout), dtype=torch.float32, device=self.dev)
        osums = torch.empty(n * 64, dtype=torch.float64, device=self.dev) if pc.cout >= 128 else None
        n_blocks = n * ((h + 15) // 16) * ((w + 15) // 16)
        if self.npass == 1 and self.use_block16 and n_blocks >= 4 * 148:
            # bf16 fast mode on the large feature maps: weight-stationary 16x16-block kernel (conv_fused3.cu), ~15 % faster there
            ops.conv3x3_fused(x, pc.hi, None, pc.cout, pc.bias, out, affine=affine, swish=swish, up2=up2, residual=residual, gn_sums=osums,
                              npass=1, block16=True)
        elif pc.f16f8 is not None:
            w16, w8pair, lo_scale = pc.f16f8
            ops.conv3x3_fused_f16f8(x, w16, w8pair, lo_scale, pc.cout, pc.bias, out, affine=affine, swish=swish, up2=up2, residual=residual,
                                    gn_sums=osums)
        else:
            ops.conv3x3_fused(x, pc.hi, pc.lo, pc.cout, pc.bias, out, affine=affine, swish=swish, up2=up2, residual=residual, gn_sums=osums,
                              npass=self.npass, two_cta=self.use_two_cta)
        if osums is not None:
            out._gn_sums = osums
        return out

    def conv3x3(self, x, name, norm=None, swish=False, residual=None, nchw=False):
        pc = self._conv(name)
        if self.use_fused and not nchw and x.shape[-1] % 64 == 0 and pc.cout % 32 == 0 and pc.ntaps == 9:
            return self._conv3x3_fused(x, pc, norm, swish, residual)
        planes = self._prep(x, norm, swish)
        return self._conv3x3_planes(planes, self._conv(name), tuple(x.shape), residual, nchw)

    def conv1x1(self, x, name, residual=None):
        n, h, w, c = x.shape
        return self._gemm_conv(self._prep(x), self._conv(name), self._TAPS1, (n, h, w, c), (n, h, w), residual)

    def resnet_block(self, x, name):
        h = self.conv3x3(x, f"{name}.conv1", norm=f"{name}.norm1", swish=True)
        sc = self.conv1x1(x, f"{name}.nin_shortcut") if self.has(f"{name}.nin_shortcut") else x
        return self.conv3x3(h, f"{name}.conv2", norm=f"{name}.norm2", swish=True, residual=sc)

    def attn_block(self, x, name):
        n, h, w, c = x.shape
        hw = h * w
        qkv_w = self._qkv(name)
        xn = self._prep(x, norm=f"{name}.norm", swish=False)
        qkv = self._gemm_conv(xn, qkv_w, self._TAPS1, (n, h, w, c), (n, h, w), out_planes=True)       # planes [n,h,w,3c]
        q_hi, q_lo = qkv
        flat = lambda t: None if t is None else t.view(n * hw, 3 * c)
        # S = q k^T  (model.py:178)
        S = torch.empty((n, hw, hw), dtype=torch.float32, device=self.dev)
        ops.gemm_tc(a_hi=q_hi, a_lo=q_lo, a_dims=(n, 1, hw, 3 * c), b_hi=flat(q_hi), b_lo=flat(q_lo), k=c, n_cols=hw,
                    b_k_off=c, b_row_zstride=hw, z_outer=n, out_w=hw, out_zo_stride=hw * hw, ldc=hw, out_f32=S,
                    bn=_bn_for(hw), npass=self.npass)
        # P = softmax(S * c^-1/2)  (:179-180); key dim padded to a multiple of 64 with zeros
        kpad = ((hw + 63) // 64) * 64
        if kpad != hw:
            p_hi = torch.zeros((n, hw, kpad), dtype=torch.bfloat16, device=self.dev)
            p_lo = torch.zeros((n, hw, kpad), dtype=torch.bfloat16, device=self.dev) if self.npass == 3 else None
        else:
            p_hi, p_lo = self._planes((n, hw, kpad))
        ops.softmax_rows(S, p_hi, p_lo, float(int(c) ** -0.5), out_ld=kpad)
        # O = P v  (:183-186), v read MN-major straight out of the qkv planes
        o_hi, o_lo = self._planes((n, h, w, c))
        ops.gemm_tc(a_hi=p_hi, a_lo=p_lo, a_dims=(n, 1, hw, kpad), b_hi=flat(q_hi), b_lo=flat(q_lo), k=kpad, n_cols=c,
                    b_k_off=2 * c, b_row_zstride=hw, z_outer=n, out_w=hw, out_zo_stride=hw * c, ldc=c, out_hi=o_hi, out_lo=o_lo,
                    flags=ops.GF_B_MN, bn=128 if c >= 128 else 64, npass=self.npass)
        return self._gemm_conv((o_hi, o_lo), self._conv(f"{name}.proj_out"), self._TAPS1, (n, h, w, c), (n, h, w), residual=x)

    def downsample(self, x, name):
        n, h, w, c = x.shape
        planes = self._prep(x, mode=ops.PREP_S2D)
        return self._gemm_conv(planes, self._conv(f"{name}.conv"), self._TAPS_S2, (n * 4, h // 2, w // 2, c), (n, h // 2, w // 2), a_n_mul=4)

    def upsample(self, x, name):
        n, h, w, c = x.shape
        pc = self._conv(f"{name}.conv")
        if self.use_fused and c % 64 == 0 and pc.cout % 32 == 0:
            return self._conv3x3_fused(x, pc, None, False, None, up2=True)
        planes = self._prep(x, mode=ops.PREP_UP2)
        return self._conv3x3_planes(planes, self._conv(f"{name}.conv"), (n, 2 * h, 2 * w, c))

    # ------------------------------------------------------------------ networks
    @torch.no_grad()
    def encoder(self, x_nchw):
        """fp32 NCHW image -> fp32 NHWC latent (N, h, w, z_channels)."""
        x_nchw = x_nchw.to(self.dev, torch.float32).contiguous()
        n, cin, H, W = x_nchw.shape
        pc_in = self._conv("encoder.conv_in", True)
        if self.use_direct_conv_in and cin == 3 and pc_in.cout in (64, 128):
            # RGB conv_in as a direct fp32 conv (exact, no im2col plane) with the GroupNorm statistics of its output fused
            if "conv_in_w32" not in self.w:
                self.w["conv_in_w32"] = self.sd["encoder.conv_in.weight"].to(self.dev, torch.float32).contiguous()
            h = torch.empty((n, H, W, pc_in.cout), dtype=torch.float32, device=self.dev)
            sums = torch.empty(n * 64, dtype=torch.float64, device=self.dev)
            ops.conv_in3(x_nchw, self.w["conv_in_w32"], pc_in.bias, h, gn_sums=sums)
            h._gn_sums = sums
        else:
            hi, lo = self._planes((n, H, W, 64))
            ops.im2col3x3(x_nchw, hi, lo)
            h = self._gemm_conv((hi, lo), pc_in, self._TAPS1, (n, H, W, 64), (n, H, W), algo_flops=2.0 * n * H * W * pc_in.cout * 9 * cin)
        for l in range(self.nlev):
            for b in range(self.nres):
                h = self.resnet_block(h, f"encoder.down.{l}.block.{b}")
                if self.has(f"encoder.down.{l}.attn.{b}.norm"):
                    h = self.attn_block(h, f"encoder.down.{l}.attn.{b}")
            if l != self.nlev - 1:
                h = self.downsample(h, f"encoder.down.{l}.downsample")
        h = self.resnet_block(h, "encoder.mid.block_1")
        h = self.attn_block(h, "encoder.mid.attn_1")
        h = self.resnet_block(h, "encoder.mid.block_2")
        return self.conv3x3(h, "encoder.conv_out", norm="encoder.norm_out", swish=True)

    @torch.no_grad()
    def decoder(self, z_nhwc):
        """fp32 NHWC latent -> fp32 NCHW image."""
        h = self.conv3x3(z_nhwc, "decoder.conv_in")
        h = self.resnet_block(h, "decoder.mid.block_1")
        h = self.attn_block(h, "decoder.mid.attn_1")
        h = self.resnet_block(h, "decoder.mid.block_2")
        for l in reversed(range(self.nlev)):
            for b in range(self.nres + 1):
                h = self.resnet_block(h, f"decoder.up.{l}.block.{b}")
                if self.has(f"decoder.up.{l}.attn.{b}.norm"):
                    h = self.attn_block(h, f"decoder.up.{l}.attn.{b}")
            if l != 0:
                h = self.upsample(h, f"decoder.up.{l}.upsample")
        pc = self._conv("decoder.conv_out")
        if self.use_direct_conv_out and pc.cout == 3 and h.shape[-1] in (64, 128):
            # RGB conv_out: GroupNorm-apply + swish + direct fp32 conv in one kernel, NHWC in -> NCHW image out (conv_small.cu)
            n, hh, ww, c = h.shape
            gamma, beta = self._norm("decoder.norm_out")
            sums = getattr(h, "_gn_sums", None)
            if sums is None:
                sums = torch.empty(n * 64, dtype=torch.float64, device=self.dev)
                mr = torch.empty(n * 64, dtype=torch.float32, device=self.dev)
                ops.groupnorm_stats(h, sums, mr, 1e-6)
            affine = torch.empty((n, c, 2), dtype=torch.float32, device=self.dev)
            ops.groupnorm_affine(sums, gamma, beta, affine, n, hh * ww, c, 1e-6)
            if "conv_out_w32" not in self.w:
                self.w["conv_out_w32"] = self.sd["decoder.conv_out.weight"].to(self.dev, torch.float32).contiguous()
            out = torch.empty((n, 3, hh, ww), dtype=torch.float32, device=self.dev)
            ops.conv_out3(h, self.w["conv_out_w32"], pc.bias, out, affine=affine, swish=True)
            return out
        if pc.cout < 16:      # image-like outputs (3 / 7 channels): the epilogue writes NCHW directly
            return self.conv3x3(h, "decoder.conv_out", norm="decoder.norm_out", swish=True, nchw=True)
        return self.nhwc_to_nchw(self.conv3x3(h, "decoder.conv_out", norm="decoder.norm_out", swish=True))

    # ------------------------------------------------------------------ VQModel-level entry points
    @torch.no_grad()
    def encode(self, x_nchw, batch=None, cam_res=None):
        """VQModel.encode -> (zq NHWC fp32 (N,h,w,e), idx int64 (N*h*w,), pre-quant h NHWC).  With `img_embed.weight` among the weights
        (geometric_embedding=True, vqgan.py:87-109) `batch` must carry intrinsics_inv / extrinsics_inv ([b, cam, 3, 3] / [b, cam, 4, 4]) and
        cam_res = (image height, image width): the ray embedding is added to the encoder output before quant_conv."""
        henc = self.encoder(x_nchw)
        if "img_embed.weight" in self.sd:
            if batch is None or cam_res is None:
                raise ValueError("this VQGAN was built with geometric_embedding=True: encode needs the batch's camera matrices and cam_res")
            n, hh, ww, d = henc.shape
            key = ("ray_pixel", hh, ww, tuple(cam_res))
            if key not in self.w:
                xs, ys = torch.linspace(0, 1, ww), torch.linspace(0, 1, hh)
                gx, gy = torch.meshgrid((xs, ys), indexing="xy")
                self.w[key] = torch.stack([gx * cam_res[1], gy * cam_res[0], torch.ones_like(gx)], -1).reshape(hh * ww, 3).to(self.dev).contiguous()
                self.w["ray_img_w"] = self.sd["img_embed.weight"].to(self.dev, torch.float32).reshape(d, 4).contiguous()
                self.w["ray_cam_w"] = self.sd["cam_embed.weight"].to(self.dev, torch.float32).reshape(d, 4).contiguous()
            I_inv = batch["intrinsics_inv"].to(self.dev, torch.float32).reshape(-1, 3, 3).contiguous()
            E_inv = batch["extrinsics_inv"].to(self.dev, torch.float32).reshape(-1, 4, 4).contiguous()
            ops.ray_embed_add(henc, I_inv, E_inv, self.w[key], self.w["ray_img_w"], self.w["ray_cam_w"])
        h = self.conv1x1(henc, "quant_conv")
        n, hh, ww, e = h.shape
        rows = n * hh * ww
        idx = torch.empty(rows, dtype=torch.int64, device=self.dev)
        zq = torch.empty((n, hh, ww, e), dtype=torch.float32, device=self.dev)
        ws = torch.empty(rows, dtype=torch.float32, device=self.dev)
        ops.vq_nearest(h.view(rows, e), self.codebook, self.code_sqnorm, ws, idx, zq.view(rows, e))
        return zq, idx, h

    @torch.no_grad()
    def codebook_entry_nhwc(self, idx, shape_bhwc):
        out = torch.empty(tuple(shape_bhwc), dtype=torch.float32, device=self.dev)
        ops.codebook_gather(self.codebook, idx.to(self.dev).reshape(-1).contiguous(), out.view(-1, shape_bhwc[-1]))
        return out

    @torch.no_grad()
    def decode_nhwc(self, zq_nhwc):
        """VQModel.decode on an NHWC latent -> fp32 NCHW reconstruction."""
        return self.decoder(self.conv1x1(zq_nhwc, "post_quant_conv"))

    @torch.no_grad()
    def decode_indices(self, idx, n, h, w):
        return self.decode_nhwc(self.codebook_entry_nhwc(idx, (n, h, w, self.embed_dim)))

    # layout helpers (C-ABI transpose kernel)
    def nhwc_to_nchw(self, x):
        n, h, w, c = x.shape
        out = torch.empty((n, c, h, w), dtype=torch.float32, device=self.dev)
        ops.transpose_f32(x, out, n, h * w, c)
        return out

    def nchw_to_nhwc(self, x):
        n, c, h, w = x.shape
        x = x.to(self.dev, torch.float32).contiguous()
        out = torch.empty((n, h, w, c), dtype=torch.float32, device=self.dev)
        ops.transpose_f32(x, out, n, c, h * w)
        return out

    # algorithmic work: conv + attention MACs*2 per image, used by bench.py
    @staticmethod
    def flops_per_image(kind="roundtrip"):
        enc, dec, small = 138.40e9, 252.72e9, 0.034e9 + 0.134e9 + 0.034e9
        return {"encode": enc + 0.168e9, "decode": dec + 0.034e9, "roundtrip": enc + dec + small}[kind]
