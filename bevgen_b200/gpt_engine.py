"""Stage-2 autoregressive multi-view transformer engine (teacher-forced forward; KV-cache sampling in gpt_decode.py).

Reference behaviour (under /root/reference/multi_view_generation): GPT.forward modules/transformer/mingpt_sparse.py:319-391,
Block.forward :240-253 (residual taken from ln1(x)), CustomSparseSelfAttention :185-212 (no output projection),
SparseSelfAttention.forward modules/transformer/sparse_self_attention.py:128-177 (bias added BEFORE the 1/sqrt(d_head) scale,
'mul' mask -> -inf), camera-bias assembly mingpt_sparse.py:375-380.

HBM layout: residual stream fp32 [B, L, d]; every GEMM consumes bf16 (hi[, lo]) operand planes written by the producing
kernel's epilogue (LayerNorm / GELU / QKV) — no separate cast passes; q/k/v live in one fused [B, L, 3d] plane pair;
weights are packed once: Wqkv [3d, d], W1 [4d, d], W2 [d, 4d], head [V, d] as K-major bf16 planes; the camera bias
(tril parameters + geometric prior) is materialised once per weight version as fp32 [L, L] together with a uint8 mask.
"""
import ctypes as C

import torch

from . import ops
from .ops import EmbedArgs


def _planes_of(w, npass, dev, min_rows=128):
    w = w.to(dev, torch.float32)
    if w.shape[0] % 128 != 0 or w.shape[0] < min_rows:
        rows = max(min_rows, ((w.shape[0] + 127) // 128) * 128)
        wp = torch.zeros(rows, w.shape[1], device=dev)
        wp[: w.shape[0]] = w
        w = wp
    return ops.split_planes(w.contiguous(), npass)


class GPTEngine:
    def __init__(self, state_dict, cfg, device="cuda", precision="fp32x3", layouts=None, use_pair_gemm=True):
        # "fp32x3": every GEMM is the bf16x3 split product; "f16f8" (parity mode): the QKV and the two MLP GEMMs form the
        # same fp32-equivalent product as 1 fp16 + 2 e4m3 MMAs (2/3 of the tensor time / energy), everything else stays bf16x3; "bf16": fast mode
        assert precision in ("fp32x3", "f16f8", "bf16")
        self.cfg, self.precision = cfg, precision
        self.npass = 1 if precision == "bf16" else 3
        self.mlp_f16f8 = precision == "f16f8"
        self.use_pair_gemm = use_pair_gemm and self.mlp_f16f8      # QKV / MLP linears on the 2-CTA 256x256 f16f8 GEMM instead of gemm_tc npass = 2
        self.dev = torch.device(device)
        dev, sd = self.dev, state_dict
        self.sd = state_dict                 # kept by reference: the KV-cache sampler packs its own 3-byte weight format from the fp32 values
        d = cfg.num_embed
        if d % 128 != 0 or d > 1024:
            raise ValueError(f"num_embed={d}: the sm_100a kernels need a multiple of 128, at most 1024")
        if cfg.hidden_size % cfg.num_heads != 0:
            raise ValueError("The hidden size (%d) is not a multiple of the number of attention heads (%d)" % (cfg.hidden_size, cfg.num_heads))
        self.d, self.H = d, cfg.num_heads
        self.dh = d // cfg.num_heads
        if self.dh != 64:
            raise ValueError(f"head size {self.dh}: the attention kernels are built for d_head = 64")
        self.L, self.nc, self.n_img, self.npad = cfg.gpt_block_size, cfg.num_cond_tokens, cfg.num_img_tokens, cfg.num_pad_tokens
        # Per-head block layouts (DeepSpeed SparsityConfig; density < 1 configs, one layout per layer: [layers][heads][nb][nb] or a single
        # [heads][nb][nb] shared by all layers).  Layouts that cover the mask (density = 1) change nothing and are dropped; otherwise
        # the fused kernel ANDs its closed-form mask with a 16-position bit table of the layout and skips empty key tiles (block sizes 16 /
        # 32 / 64 / 128); the composed attention path and the decode kernel take the uint8 layout itself.
        self.layouts = None
        if layouts is not None:
            lay = torch.as_tensor(layouts)
            lay = lay[None].expand(cfg.num_layers, *lay.shape) if lay.dim() == 3 else lay
            if lay.shape[0] != cfg.num_layers:
                raise ValueError(f"expected {cfg.num_layers} per-layer layouts, got {lay.shape[0]}")
            if not all(cfg.layout_covers_mask(l) for l in lay):
                self.layouts = (lay != 0).to(self.dev, torch.uint8).contiguous()
        self.layout_block = cfg.sparse_block_size
        f32 = lambda k: sd[k].detach().to(dev, torch.float32).contiguous()
        self.layers = []
        for i in range(cfg.num_layers):
            p = f"blocks.{i}"
            wqkv = torch.cat([sd[f"{p}.attention.{n}.weight"] for n in ("query", "key", "value")], 0)
            bqkv = torch.cat([sd[f"{p}.attention.{n}.bias"] for n in ("query", "key", "value")], 0)
            self.layers.append(dict(
                ln1=(f32(f"{p}.ln1.weight"), f32(f"{p}.ln1.bias")), ln2=(f32(f"{p}.ln2.weight"), f32(f"{p}.ln2.bias")),
                wqkv=_planes_of(wqkv, self.npass, dev), bqkv=bqkv.to(dev, torch.float32).contiguous(),
                w1=_planes_of(sd[f"{p}.mlp.0.weight"], self.npass, dev), b1=f32(f"{p}.mlp.0.bias"),
                w2=_planes_of(sd[f"{p}.mlp.2.weight"], self.npass, dev), b2=f32(f"{p}.mlp.2.bias")))
            if self.mlp_f16f8:       # (fp16 plane, e4m3 pair plane, lo_scale) of the QKV / MLP weights for the npass = 2 GEMM
                self.layers[-1]["wqkv_f8"] = ops.pack_f16f8(wqkv.to(dev, torch.float32).contiguous())
                self.layers[-1]["w1_f8"] = ops.pack_f16f8(sd[f"{p}.mlp.0.weight"].to(dev, torch.float32).contiguous())
                self.layers[-1]["w2_f8"] = ops.pack_f16f8(sd[f"{p}.mlp.2.weight"].to(dev, torch.float32).contiguous())
                if self.use_pair_gemm:   # scaled single-accumulator operands of the 2-CTA GEMM (gemm_pair.cu)
                    self.layers[-1]["wqkv_p"] = ops.pack_linear_f16f8(wqkv.to(dev, torch.float32).contiguous())
                    self.layers[-1]["w1_p"] = ops.pack_linear_f16f8(sd[f"{p}.mlp.0.weight"].to(dev, torch.float32).contiguous())
                    self.layers[-1]["w2_p"] = ops.pack_linear_f16f8(sd[f"{p}.mlp.2.weight"].to(dev, torch.float32).contiguous())
        self.ln_f = (f32("ln_f.weight"), f32("ln_f.bias"))
        self.vocab = sd["head.weight"].shape[0]
        self.whead = _planes_of(sd["head.weight"], self.npass, dev)
        # embeddings
        self.x_tok_emb, self.cond_tok_emb = f32("x_tok_emb.weight"), f32("cond_tok_emb.weight")
        self.x_pos_emb = f32("x_pos_emb").reshape(-1, d).contiguous()
        cond_static = sd["cond_pos_emb"].detach().to(dev, torch.float32).reshape(-1, d).clone()
        self.image_embed = "img_embed.weight" in sd and cfg.image_embed
        self.bev_embed = "bev_embed.weight" in sd and cfg.bev_embed
        self.img_w = f32("img_embed.weight").reshape(d, 4).contiguous() if self.image_embed else None
        self.cam_w = f32("cam_embed.weight").reshape(d, 4).contiguous() if self.image_embed else None
        if self.bev_embed:      # weight-only part of mingpt_sparse.py:352-358 folded at pack time
            from .geometry_torch import bev_grid
            g = bev_grid(*cfg.bev_latent_res)[:2].reshape(2, -1).to(dev)
            grid_embed = (f32("bev_embed.weight").reshape(d, 2) @ g).t() + f32("bev_embed.bias")
            cond_static = cond_static + grid_embed - f32("bev_cam_pos_emb")[0].sum(0)
        self.cond_static = cond_static.contiguous()
        from .geometry_torch import image_plane
        self.pixel = image_plane(cfg.cam_latent_h, cfg.cam_latent_w, cfg.cam_res).to(dev).contiguous()     # [hw, 3]
        self.fwd = cfg.forward_shuffle_idx.to(dev, torch.int32).contiguous()
        self.bwd = cfg.backward_shuffle_idx.to(dev)
        # camera bias + mask, once per weight version (the reference rebuilds them every forward)
        self.mask_u8 = (cfg.attention_mask != 0).to(dev, torch.uint8).contiguous()
        self.bias = None
        if cfg.camera_bias and "camera_bias_emb" in sd:
            L = self.L
            idx = torch.tril_indices(L, L, device=dev)
            b = torch.zeros(L, L, dtype=torch.float32, device=dev)
            b[idx[0], idx[1]] = f32("camera_bias_emb")[0]
            self.bias = (b + cfg.prob_matrix.to(dev, torch.float32)).contiguous()
        i, j = torch.meshgrid(torch.arange(self.L), torch.arange(self.L), indexing="ij")
        closed = (j < self.nc) | ((i >= self.nc) & (j <= i))
        self.causal = bool(torch.equal(closed, cfg.attention_mask.bool() | closed) and self.npad == 0)   # mask ⊆ [cond | causal]
        # KV-cache decoding only needs the closed form on the REAL rows / columns: trailing pad tokens (non-square latents padded to the
        # block size, mask_generator.py:197-205) are never queried and never visible to a real row, so they are simply ignored
        real = self.nc + self.n_img
        am = cfg.attention_mask.bool()
        self.decode_causal = bool(torch.equal(am[:real, :real], closed[:real, :real]) and not am[:real, real:].any())
        # Padded geometries (non-square latents: nuScenes-native 14x25 -> L = 2368 with 12 pad tokens, mask_generator.py:197-205; and any
        # L that is not a multiple of the 128-row attention tile): the forward runs on Lrun = roundup(L, 128) rows per sample.  Rows
        # beyond the real tokens (the reference's pad tokens + our extension rows) carry the PAD embedding; no real row ever attends to
        # them (the closed form [cond | causal] holds on the real rows), their own outputs are never read (the reference slices them
        # off, mingpt_sparse.py:387), so the fused kernel's closed-form mask applies unchanged and S is never materialised.
        self.Lrun = self.L
        self.fused_pad = bool(self.decode_causal and not self.causal and self.nc % 128 == 0)
        if self.fused_pad:
            self.Lrun = ((self.L + 127) // 128) * 128
        # tiled, pre-scaled fp16 copy for the fused kernel (coalesced 16-byte reads per lane, scale folded into one FFMA per score)
        self.bias_f16 = None
        if self.bias is not None and self.Lrun % 128 == 0:
            bpad = self.bias
            if self.Lrun != self.L:
                bpad = torch.zeros((self.Lrun, self.Lrun), dtype=torch.float32, device=dev)
                bpad[: self.L, : self.L] = self.bias
            self.bias_f16 = ops.tile_attention_bias(bpad, float(self.dh) ** -0.5)
        self.fused_attention = True          # tcgen05 flash-style kernel when the geometry allows; composed path otherwise
        self._perm_cache = {}
        self._allowed = float(self.mask_u8.sum().item())      # attended (row, col) pairs: algorithmic attention work
        if self.layouts is not None:
            fusable = self.layout_block in (16, 32, 64, 128) and self.L % 128 == 0 and self.L <= 4096
            for i, lw in enumerate(self.layers):
                lw["layout"] = self.layouts[i]
                # the fused kernel's 16-position bit table: key tiles without any block of the head's layout are skipped
                lw["layout64"] = ops.layout_to_tiles64(self.layouts[i], self.layout_block, self.L) if fusable else None

    # ------------------------------------------------------------------ helpers
    def _planes(self, shape):
        hi = torch.empty(shape, dtype=torch.bfloat16, device=self.dev)
        lo = torch.empty(shape, dtype=torch.bfloat16, device=self.dev) if self.npass == 3 else None
        return hi, lo

    def _linear(self, a, w, n_cols, rows, k, bias=None, residual=None, out_f32=None, out_planes=None, flags=0, taps=((0, 0, 0),),
                a_dims=None, out_w=None, z_outer=1, out_zo_stride=0, npass=None, lo_scale=0.0):
        a_hi, a_lo = a
        oh, ol = out_planes if out_planes is not None else (None, None)
        ops.gemm_tc(a_hi=a_hi, a_lo=a_lo, a_dims=a_dims or (1, 1, rows, k), b_hi=w[0], b_lo=w[1], k=k, n_cols=n_cols, taps=taps,
                    out_w=out_w or rows, z_outer=z_outer, out_zo_stride=out_zo_stride, ldc=n_cols, bias=bias, residual=residual,
                    out_f32=out_f32, out_hi=oh, out_lo=ol, flags=flags, bn=128, npass=self.npass if npass is None else npass, lo_scale=lo_scale)

    def embed(self, cam_idx, bev_idx, batch, sampling, row0=0, nrows=None, out=None):
        B = bev_idx.shape[0]
        nrows = self.L - row0 if nrows is None else nrows
        out = torch.empty((B, nrows, self.d), dtype=torch.float32, device=self.dev) if out is None else out
        a = EmbedArgs()
        keep = [cam_idx.to(self.dev, torch.int64).contiguous(), bev_idx.to(self.dev, torch.int64).contiguous(),
                batch["intrinsics_inv"].to(self.dev, torch.float32).contiguous(), batch["extrinsics_inv"].to(self.dev, torch.float32).contiguous()]
        a.cam_idx, a.bev_idx, a.intrinsics_inv, a.extrinsics_inv = (t.data_ptr() for t in keep)
        a.x_tok_emb, a.cond_tok_emb = self.x_tok_emb.data_ptr(), self.cond_tok_emb.data_ptr()
        a.x_pos_emb, a.cond_static = self.x_pos_emb.data_ptr(), self.cond_static.data_ptr()
        a.img_embed_w = self.img_w.data_ptr() if self.image_embed else None
        a.cam_embed_w = self.cam_w.data_ptr() if self.image_embed else None
        a.forward_shuffle_idx, a.pixel, a.out = self.fwd.data_ptr(), self.pixel.data_ptr(), out.data_ptr()
        a.B, a.ncam, a.hw, a.nc, a.n_img, a.L, a.d, a.vocab = B, self.cfg.num_cams, self.cfg.num_cam_tokens, self.nc, self.n_img, max(self.L, row0 + nrows), self.d, self.cfg.vocab_size
        a.pad_last, a.bev_embed, a.row0, a.nrows = int(not sampling), int(self.bev_embed), row0, nrows
        ops.embed_assemble(a)
        return out

    def attention(self, qkv, y, B, L, bias="full", mask=None, causal=None, allowed=None, fused_cond=None, layout=None, layout64=None):
        """qkv planes [B, L, 3d]; returns x1 = y + concat_heads(softmax(scale*(QK^T + bias))V) as fp32 [B, L, d].
        bias/mask default to the full-sequence camera bias and attention mask; the KV-cache prefill passes the cond x cond blocks."""
        d, H, dh = self.d, self.H, self.dh
        if layout is not None:
            fused_cond = None
        if (self.fused_attention and (layout is None or layout64 is not None) and isinstance(bias, str) and mask is None
                and (self.causal or (self.fused_pad and layout is None)) and L % 128 == 0 and self.nc % 128 == 0):
            x1 = torch.empty((B, L, d), dtype=torch.float32, device=self.dev)
            ops.attn_fused_fwd(qkv[0], qkv[1], B, L, H, d, self.nc, self.bias_f16, y, x1, float(dh) ** -0.5, self.npass,
                               algo_flops=4.0 * B * H * dh * self._allowed, layout64=layout64 if layout is not None else None)
            return x1
        if self.fused_attention and fused_cond is not None and L % 128 == 0:
            x1 = torch.empty((B, L, d), dtype=torch.float32, device=self.dev)
            ops.attn_fused_fwd(qkv[0], qkv[1], B, L, H, d, L, fused_cond, y, x1, float(dh) ** -0.5, self.npass,
                               algo_flops=4.0 * B * H * dh * float(L * L))
            return x1
        bias = self.bias if isinstance(bias, str) else bias
        mask = self.mask_u8 if mask is None else mask
        causal = self.causal if causal is None else causal
        allowed = self._allowed if allowed is None else allowed
        q_hi, q_lo = qkv
        flat = lambda t: None if t is None else t.view(B * L, 3 * d)
        S = torch.empty((B, H, L, L), dtype=torch.float32, device=self.dev)
        cz = ops.GF_CAUSAL_SKIP if causal else 0
        ops.gemm_tc(a_hi=q_hi, a_lo=q_lo, a_dims=(B, 1, L, 3 * d), b_hi=flat(q_hi), b_lo=flat(q_lo), k=dh, n_cols=L, a_c_zstride=dh,
                    b_k_off=d, b_k_zstride=dh, b_row_zstride=L, z_inner=H, z_outer=B, out_w=L, out_zo_stride=H * L * L,
                    out_zi_stride=L * L, ldc=L, out_f32=S, flags=cz, causal_ncond=self.nc, bn=128, npass=self.npass,
                    algo_flops=2.0 * B * H * dh * allowed)
        p_hi, p_lo = self._planes((B, H, L, L))
        ops.attn_softmax(S, bias, mask, p_hi, p_lo, L, float(dh) ** -0.5, layout=layout, heads=H, block=self.layout_block)
        x1 = torch.empty((B, L, d), dtype=torch.float32, device=self.dev)
        kz = ops.GF_CAUSAL_KLIMIT if causal else 0
        ops.gemm_tc(a_hi=p_hi, a_lo=p_lo, a_dims=(B * H, 1, L, L), b_hi=flat(q_hi), b_lo=flat(q_lo), k=L, n_cols=dh, a_n_mul=H, a_n_zstride=1,
                    b_k_off=2 * d, b_k_zstride=dh, b_row_zstride=L, z_inner=H, z_outer=B, out_w=L, out_zo_stride=L * d, out_zi_stride=dh,
                    ldc=d, residual=y, out_f32=x1, flags=ops.GF_B_MN | kz, causal_ncond=self.nc, bn=64, npass=self.npass,
                    algo_flops=2.0 * B * H * dh * allowed)
        return x1

    def block(self, x, lw, B, L, attn_kw=None, on_qkv=None):
        d = self.d
        rows = B * L
        y = torch.empty_like(x)
        qkv = self._planes((B, L, 3 * d))
        pair = self.use_pair_gemm and d % 128 == 0
        if self.mlp_f16f8 and d % 128 == 0:
            yp = (torch.empty((rows, d), dtype=torch.float16, device=self.dev), torch.empty((rows, 2 * d), dtype=torch.uint8, device=self.dev))
            ops.layernorm(x, *lw["ln1"], y=y, out_hi=yp[0], out_lo=yp[1], f16f8=True, scaled=pair)
            if pair:
                wq = lw["wqkv_p"]
                ops.linear_f16f8(yp[0], yp[1], wq[0], wq[1], wq[2], rows, 3 * d, d, bias=lw["bqkv"], out_hi=qkv[0], out_lo=qkv[1])
            else:
                wq = lw["wqkv_f8"]
                self._linear(yp, wq[:2], 3 * d, rows, d, bias=lw["bqkv"], out_planes=qkv, npass=2, lo_scale=wq[2])     # output planes stay bf16 hi/lo
        else:
            yp = self._planes((rows, d))
            ops.layernorm(x, *lw["ln1"], y=y, out_hi=yp[0], out_lo=yp[1])
            self._linear(yp, lw["wqkv"], 3 * d, rows, d, bias=lw["bqkv"], out_planes=qkv)
        if on_qkv is not None:
            on_qkv(qkv)
        akw = dict(attn_kw or {})
        if "layout" not in akw and lw.get("layout") is not None:
            akw["layout"] = lw["layout"]
            akw["layout64"] = lw.get("layout64")
        x1 = self.attention(qkv, y, B, L, **akw)
        x2 = torch.empty_like(x)
        if self.mlp_f16f8 and d % 128 == 0:
            # MLP as f16f8 GEMMs: LayerNorm and the GELU epilogue write the fp16 + e4m3-pair operand planes directly
            u8 = lambda r, c: torch.empty((r, c), dtype=torch.uint8, device=self.dev)
            f16 = lambda r, c: torch.empty((r, c), dtype=torch.float16, device=self.dev)
            zp, hp = (f16(rows, d), u8(rows, 2 * d)), (f16(rows, 4 * d), u8(rows, 8 * d))
            ops.layernorm(x1, *lw["ln2"], out_hi=zp[0], out_lo=zp[1], f16f8=True, scaled=pair)
            if pair:
                w1, w2 = lw["w1_p"], lw["w2_p"]
                ops.linear_f16f8(zp[0], zp[1], w1[0], w1[1], w1[2], rows, 4 * d, d, bias=lw["b1"], gelu=True, out_f16=hp[0], out_pair=hp[1])
                ops.linear_f16f8(hp[0], hp[1], w2[0], w2[1], w2[2], rows, d, 4 * d, bias=lw["b2"], residual=x1, out_f32=x2)
                return x2
            w1, w2 = lw["w1_f8"], lw["w2_f8"]
            self._linear(zp, w1[:2], 4 * d, rows, d, bias=lw["b1"], out_planes=hp, flags=ops.GF_GELU | ops.GF_OUT_F16F8, npass=2, lo_scale=w1[2])
            self._linear(hp, w2[:2], d, rows, 4 * d, bias=lw["b2"], residual=x1, out_f32=x2, npass=2, lo_scale=w2[2])
            return x2
        zp = self._planes((rows, d))
        ops.layernorm(x1, *lw["ln2"], out_hi=zp[0], out_lo=zp[1])
        hp = self._planes((rows, 4 * d))
        self._linear(zp, lw["w1"], 4 * d, rows, d, bias=lw["b1"], out_planes=hp, flags=ops.GF_GELU)
        self._linear(hp, lw["w2"], d, rows, 4 * d, bias=lw["b2"], residual=x1, out_f32=x2)
        return x2

    def _perm(self, B):
        if B not in self._perm_cache:
            base = torch.arange(B, device=self.dev)[:, None] * self.n_img
            self._perm_cache[B] = (base + self.bwd[None, :]).reshape(-1).contiguous()
        return self._perm_cache[B]

    @torch.no_grad()
    def forward(self, cam_idx, bev_idx, batch, sampling, return_hidden=False):
        """-> logits fp32 [B, n_img, vocab] in (cam, h, w) order (GPT.forward)."""
        B, d = bev_idx.shape[0], self.d
        L = self.Lrun if (self.fused_attention and self.layouts is None) else self.L      # padded geometries: see __init__
        x = self.embed(cam_idx, bev_idx, batch, sampling, nrows=L)
        hidden = []
        for lw in self.layers:
            x = self.block(x, lw, B, L)
            if return_hidden:
                hidden.append(x[:, : self.L])
        fp = self._planes((B, L, d))
        ops.layernorm(x, *self.ln_f, out_hi=fp[0], out_lo=fp[1])
        # head on rows n_cond-1 .. n_cond+n_img-2 only (position p predicts token p+1, :390): a row-shifted 1-tap GEMM
        dec = torch.empty((B, self.n_img, self.vocab), dtype=torch.float32, device=self.dev)
        self._linear(fp, self.whead, self.vocab, None, d, out_f32=dec, taps=((self.nc - 1, 0, 0),), a_dims=(B, 1, L, d),
                     out_w=self.n_img, z_outer=B, out_zo_stride=self.n_img * self.vocab)
        logits = torch.empty_like(dec)
        ops.codebook_gather(dec.view(B * self.n_img, self.vocab), self._perm(B), logits.view(B * self.n_img, self.vocab))
        return (logits, hidden) if return_hidden else logits

    def flops_per_sample(self):
        """Algorithmic FLOPs of one teacher-forced forward: linears + allowed-only attention + head."""
        L, d = self.L, self.d
        lin = self.cfg.num_layers * 22 * L * d * d
        att = self.cfg.num_layers * 4 * self._allowed * d
        head = 2 * self.n_img * d * self.vocab
        return lin + att + head
