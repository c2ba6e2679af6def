"""MaskGit stage-2 variant on the CUDA library: the bidirectional decoder the released weights use.

Replaces, for inference, modules/stage2/muse_maskgit_pytorch.py of the reference:
  TransformerMultiView.forward :283-366, Attention :90-169 (self + cross, null key/value, cosine-sim q/k, additive camera bias slices),
  FeedForward/GEGLU :72-88, TransformerBlocks :171-202, SelfCritic :371-381, MaskGit.generate :511-627.
Every Linear and both attention products run on `bevgen_gemm_tc` (bf16x3 split products in the parity mode), LayerNorm / the biased
softmax / the input embeddings on the stage-2 kernels shared with the autoregressive decoder, and the operand planes between them on
`bevgen_mg_head_planes` / `bevgen_mg_geglu_ln`.  No CPU or eager fallback: a missing library raises.

Classifier-free guidance: in eval mode the reference's conditioning dropout is inactive (:341 `if self.training and ...`), so the "null"
forward of `forward_with_cond_scale` is bit-identical to the conditional forward and the guided logits equal the plain logits.  The
engine therefore runs ONE forward per de-masking step (and one for the critic) where the reference runs two each.
"""
import math

import torch

from . import ops
from ._lib import EmbedArgs
from .gpt_engine import _planes_of


class MaskGitEngine:
    def __init__(self, state_dict, cfg, depth, heads, dim_head=64, ff_mult=4, device="cuda", precision="fp32x3", critic=None):
        # "fp32x3": every GEMM is the bf16x3 split product on gemm_tc; "f16f8" (default parity mode): the large Linear layers (self q|k|v,
        # cross q, both feed-forward layers = 86 % of the linear FLOPs) run on the 2-CTA f16f8 GEMM (gemm_pair.cu), the rest stays bf16x3
        assert precision in ("fp32x3", "f16f8", "bf16")
        if dim_head != 64:
            raise ValueError("the attention kernels are built for d_head = 64")
        if cfg.num_pad_tokens != 0:
            raise ValueError("MaskGit needs gpt_block_size == num_cond_tokens + num_img_tokens (the reference slices the camera bias "
                             "with that assumption, muse_maskgit_pytorch.py:150-156); use sparse_block_size=1 as its config does")
        self.cfg, self.precision = cfg, precision
        self.npass = 1 if precision == "bf16" else 3
        self.pair = precision == "f16f8"
        self.dev = dev = torch.device(device)
        sd = state_dict
        self.d = d = cfg.num_embed
        self.H, self.inner = heads, heads * dim_head
        self.depth = depth
        self.nc, self.n_img, self.L = cfg.num_cond_tokens, cfg.num_img_tokens, cfg.gpt_block_size
        self.f = int(d * ff_mult * 2 / 3)
        self.f_pad = -(-self.f // 64) * 64
        self.n1_pad = -(-2 * self.f // 32) * 32
        if d % 128 or d > 1024 or self.f_pad > 3072:
            raise ValueError(f"unsupported width {d}")
        self.vocab = sd["to_logits.weight"].shape[0]
        self.mask_id = sd["token_emb.weight"].shape[0] - 1
        f32 = lambda k: sd[k].detach().to(dev, torch.float32).contiguous()
        pl = lambda w: _planes_of(w.detach().to(dev, torch.float32), self.npass, dev)
        self.zero_d = torch.zeros(d, device=dev)
        self.layers = []
        for i in range(depth):
            p = f"transformer_blocks.layers.{i}"
            lw = {}
            for a, name in ((0, "self"), (1, "cross")):
                q = f"{p}.{a}"
                null = f32(f"{q}.null_kv")                                    # [2][H][1][64]
                lw[name] = dict(gamma=f32(f"{q}.norm.gamma"), q_scale=f32(f"{q}.q_scale"), k_scale=f32(f"{q}.k_scale"),
                                null_k=null[0].reshape(heads, 64).contiguous(), null_v=null[1].reshape(heads, 64).contiguous(),
                                wout=pl(sd[f"{q}.to_out.weight"]))
                if a == 0:      # one GEMM for q | k | v of the self-attention
                    lw[name]["wqkv"] = pl(torch.cat([sd[f"{q}.to_q.weight"], sd[f"{q}.to_kv.weight"]], 0))
                else:
                    lw[name]["wq"], lw[name]["wkv"] = pl(sd[f"{q}.to_q.weight"]), pl(sd[f"{q}.to_kv.weight"])
            w2 = torch.zeros(d, self.f_pad, device=dev)
            w2[:, : self.f] = sd[f"{p}.2.4.weight"].detach().to(dev, torch.float32)
            lw["ff"] = dict(gamma0=f32(f"{p}.2.0.gamma"), w1=pl(sd[f"{p}.2.1.weight"]), gamma3=f32(f"{p}.2.3.gamma"), w2=pl(w2))
            if self.pair:
                pk = lambda w: ops.pack_linear_f16f8(w.detach().to(dev, torch.float32).contiguous())
                lw["self"]["wqkv_p"] = pk(torch.cat([sd[f"{p}.0.to_q.weight"], sd[f"{p}.0.to_kv.weight"]], 0))
                lw["cross"]["wq_p"] = pk(sd[f"{p}.1.to_q.weight"])
                w1 = torch.zeros(self.n1_pad, d, device=dev)            # output features padded to the GEMM's 32-column epilogue chunk
                w1[: 2 * self.f] = sd[f"{p}.2.1.weight"].detach().to(dev, torch.float32)
                lw["ff"]["w1_p"], lw["ff"]["w2_p"] = pk(w1), pk(w2)
            self.layers.append(lw)
        self.gamma_f = f32("transformer_blocks.norm.gamma")
        self.w_logits = pl(sd["to_logits.weight"])
        self.critic = None
        if critic is not None:      # SelfCritic.to_pred: Linear(d, 1) on the final embeddings
            self.critic = (pl(critic["weight"]), critic["bias"].detach().to(dev, torch.float32).contiguous())
        # embeddings (same assembly kernel as the autoregressive decoder; identity token order, no pad rows)
        self.tok_emb, self.cond_tok_emb = f32("token_emb.weight"), f32("cond_token_emb.weight")
        self.pos_emb = f32("pos_emb.weight")
        cond_static = f32("cond_pos_emb.weight").clone()
        self.image_embed = bool(cfg.image_embed) and "img_embed.weight" in sd
        self.bev_embed = bool(cfg.bev_embed) and "bev_embed.weight" in sd
        self.img_w = f32("img_embed.weight").reshape(d, 4).contiguous() if self.image_embed else None
        self.cam_w = f32("cam_embed.weight").reshape(d, 4).contiguous() if self.image_embed else None
        if self.bev_embed:
            from .geometry_torch import bev_grid
            g = bev_grid(*cfg.bev_latent_res)[:2].reshape(2, -1).to(dev)
            cond_static = cond_static + ((f32("bev_embed.weight").reshape(d, 2) @ g).t() + f32("bev_embed.bias")) - f32("bev_cam_pos_emb")[0].sum(0)
        self.cond_static = cond_static.contiguous()
        from .geometry_torch import image_plane
        self.pixel = image_plane(cfg.cam_latent_h, cfg.cam_latent_w, cfg.cam_res).to(dev).contiguous()
        self.order = torch.arange(self.n_img, dtype=torch.int32, device=dev)
        # camera bias slices with the null column, divided by the similarity scale (the softmax kernel computes scale * (s + bias)), padded
        # to the key count of the GEMM tiles; the mask switches the padding keys off
        self.scale = 8.0
        self.lk_self, self.lk_cross = -(-(self.n_img + 1) // 128) * 128, -(-(self.nc + 1) // 128) * 128
        bias = None
        if cfg.camera_bias and "camera_bias_emb" in sd:
            idx = torch.tril_indices(self.L, self.L, device=dev)
            bias = torch.zeros(self.L, self.L, device=dev)
            bias[idx[0], idx[1]] = f32("camera_bias_emb")[0]
            bias = bias + cfg.prob_matrix.to(dev, torch.float32)

        def table(sl, n_keys, lk):
            t = torch.zeros(self.n_img, lk, device=dev)
            if sl is not None:
                t[:, 1: n_keys + 1] = sl / self.scale
            m = torch.zeros(self.n_img, lk, dtype=torch.uint8, device=dev)
            m[:, : n_keys + 1] = 1
            return t.contiguous(), m.contiguous()
        self.bias_self, self.mask_self = table(None if bias is None else bias[self.nc:, self.nc:], self.n_img, self.lk_self)
        self.bias_cross, self.mask_cross = table(None if bias is None else bias[self.nc:, : self.nc], self.nc, self.lk_cross)
        # self-attention on the fused flash-style kernel: q | k | v share one plane of lk_self rows per scene (query i at row i, null key at
        # row 0, key j at row j + 1), dense support (n_cond = seq_len), padding keys switched off by -inf entries of the tiled bias table
        self.lk_f = max(self.lk_self, self.lk_cross)          # rows per scene of the fused kernel's shared q | k | v plane
        self.fused_self = self.lk_f <= 4096
        if self.fused_self:
            nt = self.lk_f // 128

            def fused_tables(tab, n_keys, lk):
                full = torch.zeros(self.lk_f, self.lk_f, device=dev)
                full[: self.n_img, :lk] = tab
                full[:, n_keys + 1:] = float("-inf")
                t = torch.zeros(self.H, nt, nt, dtype=torch.int64, device=dev)
                t[:, :, : lk // 128] = -1
                return ops.tile_attention_bias(full, self.scale), (None if lk == self.lk_f else t.contiguous())
            self.bias_self_tiled, self.self_tiles = fused_tables(self.bias_self, self.n_img, self.lk_self)
            # cross-attention on the same kernel and geometry: the context keys (null + n_cond, padded to lk_cross) occupy the first key
            # tiles, the layout table (SURVEY 8f-2 mechanism) makes every role skip the remaining ones
            self.bias_cross_tiled, self.cross_tiles = fused_tables(self.bias_cross, self.nc, self.lk_cross)

    # ------------------------------------------------------------------ helpers
    def _planes(self, shape):
        hi = torch.empty(shape, dtype=torch.bfloat16, device=self.dev)
        lo = torch.empty(shape, dtype=torch.bfloat16, device=self.dev) if self.npass == 3 else None
        return hi, lo

    def _linear(self, a, w, n_cols, rows, k, bias=None, residual=None, out_f32=None, out_planes=None):
        oh, ol = out_planes if out_planes is not None else (None, None)
        ops.gemm_tc(a_hi=a[0], a_lo=a[1], a_dims=(1, 1, rows, k), b_hi=w[0], b_lo=w[1], k=k, n_cols=n_cols, out_w=rows, ldc=n_cols, bias=bias,
                    residual=residual, out_f32=out_f32, out_hi=oh, out_lo=ol, bn=128, npass=self.npass)

    def _ln_planes(self, x, gamma, rows, y=None, f16f8=False):
        if f16f8:       # scaled fp16 plane + e4m3 pair plane: A operand of bevgen_linear_f16f8
            yp = (torch.empty((rows, self.d), dtype=torch.float16, device=self.dev), torch.empty((rows, 2 * self.d), dtype=torch.uint8, device=self.dev))
            ops.layernorm(x, gamma, self.zero_d, y=y, out_hi=yp[0], out_lo=yp[1], rows=rows, f16f8=True, scaled=True)
            return yp
        yp = self._planes((rows, self.d))
        ops.layernorm(x, gamma, self.zero_d, y=y, out_hi=yp[0], out_lo=yp[1], rows=rows)
        return yp

    def _pair_linear(self, a, w, n_cols, rows, k, residual=None, out_f32=None):
        ops.linear_f16f8(a[0], a[1], w[0], w[1], w[2], rows, n_cols, k, residual=residual, out_f32=out_f32)

    def _attend(self, q_src, q_ld, q_col0, kv_src, kv_ld, k_col0, v_col0, n_kv, lk, aw, bias, mask, B, residual):
        """softmax(8 * l2norm(q).l2norm(k) + bias) v over [null | n_kv keys], then to_out + residual -> fp32 [B*n_img, d]."""
        n, H, inner = self.n_img, self.H, self.inner
        qp, kp, vp = self._planes((B * n, inner)), self._planes((B * lk, inner)), self._planes((B * lk, inner))
        ops.mg_head_planes(q_src, q_ld, q_col0, n, qp[0], qp[1], B, n, H, scale=aw["q_scale"])
        ops.mg_head_planes(kv_src, kv_ld, k_col0, n_kv, kp[0], kp[1], B, lk, H, null_vec=aw["null_k"], scale=aw["k_scale"])
        ops.mg_head_planes(kv_src, kv_ld, v_col0, n_kv, vp[0], vp[1], B, lk, H, null_vec=aw["null_v"])
        S = torch.empty((B, H, n, lk), dtype=torch.float32, device=self.dev)
        ops.gemm_tc(a_hi=qp[0], a_lo=qp[1], a_dims=(B, 1, n, inner), b_hi=kp[0], b_lo=kp[1], k=64, n_cols=lk, a_c_zstride=64, b_k_zstride=64,
                    b_row_zstride=lk, z_inner=H, z_outer=B, out_w=n, out_zo_stride=H * n * lk, out_zi_stride=n * lk, ldc=lk, out_f32=S,
                    bn=128, npass=self.npass)
        pp = self._planes((B, H, n, lk))
        ops.attn_softmax(S, bias, mask, pp[0], pp[1], n, self.scale)
        del S
        op = self._planes((B * n, inner))
        ops.gemm_tc(a_hi=pp[0], a_lo=pp[1], a_dims=(B * H, 1, n, lk), b_hi=vp[0], b_lo=vp[1], k=lk, n_cols=64, a_n_mul=H, a_n_zstride=1,
                    b_k_zstride=64, b_row_zstride=lk, z_inner=H, z_outer=B, out_w=n, out_zo_stride=n * inner, out_zi_stride=64, ldc=inner,
                    out_hi=op[0], out_lo=op[1], flags=ops.GF_B_MN, bn=64, npass=self.npass)
        out = torch.empty((B * n, self.d), dtype=torch.float32, device=self.dev)
        self._linear(op, aw["wout"], self.d, B * n, inner, residual=residual, out_f32=out)
        return out

    def _attend_self_fused(self, qkv, aw, B, residual):
        """Self-attention through bevgen_attn_fused_fwd (no score matrix in HBM): operand planes [B][lk][3*inner], output planes -> to_out."""
        n, H, inner, lk = self.n_img, self.H, self.inner, self.lk_f
        fp = self._planes((B * lk, 3 * inner))
        ops.mg_head_planes(qkv, 3 * inner, 0, n, fp[0], fp[1], B, lk, H, scale=aw["q_scale"], dst_ld=3 * inner, dst_col0=0)
        ops.mg_head_planes(qkv, 3 * inner, inner, n, fp[0], fp[1], B, lk, H, null_vec=aw["null_k"], scale=aw["k_scale"], dst_ld=3 * inner, dst_col0=inner)
        ops.mg_head_planes(qkv, 3 * inner, 2 * inner, n, fp[0], fp[1], B, lk, H, null_vec=aw["null_v"], dst_ld=3 * inner, dst_col0=2 * inner)
        op = self._planes((B * lk, inner))
        ops.attn_fused_fwd(fp[0], fp[1], B, lk, H, inner, lk, self.bias_self_tiled, None, None, self.scale, self.npass,
                           algo_flops=4.0 * B * H * 64 * float(n) * (n + 1), layout64=self.self_tiles, out_hi=op[0], out_lo=op[1])
        out = torch.empty((B * n, self.d), dtype=torch.float32, device=self.dev)
        # rows n .. lk-1 of every scene are padding queries: the GEMM reads the [B][lk] plane and stores the first n rows of each scene
        ops.gemm_tc(a_hi=op[0], a_lo=op[1], a_dims=(B, 1, lk, inner), b_hi=aw["wout"][0], b_lo=aw["wout"][1], k=inner, n_cols=self.d, out_w=n,
                    z_outer=B, out_zo_stride=n * self.d, ldc=self.d, residual=residual, out_f32=out, bn=128, npass=self.npass)
        return out

    def _attend_cross_fused(self, q, kv, aw, B, residual):
        n, H, inner, lk = self.n_img, self.H, self.inner, self.lk_f
        fp = self._planes((B * lk, 3 * inner))
        ops.mg_head_planes(q, inner, 0, n, fp[0], fp[1], B, lk, H, scale=aw["q_scale"], dst_ld=3 * inner, dst_col0=0)
        # only the first lk_cross key rows of each scene are ever loaded (the layout table skips the other key tiles)
        kr = self.lk_cross
        ops.mg_head_planes(kv, 2 * inner, 0, self.nc, fp[0], fp[1], B, kr, H, null_vec=aw["null_k"], scale=aw["k_scale"], dst_ld=3 * inner, dst_col0=inner,
                           dst_batch_rows=lk)
        ops.mg_head_planes(kv, 2 * inner, inner, self.nc, fp[0], fp[1], B, kr, H, null_vec=aw["null_v"], dst_ld=3 * inner, dst_col0=2 * inner,
                           dst_batch_rows=lk)
        op = self._planes((B * lk, inner))
        ops.attn_fused_fwd(fp[0], fp[1], B, lk, H, inner, lk, self.bias_cross_tiled, None, None, self.scale, self.npass,
                           algo_flops=4.0 * B * H * 64 * float(n) * (self.nc + 1), layout64=self.cross_tiles, out_hi=op[0], out_lo=op[1])
        out = torch.empty((B * n, self.d), dtype=torch.float32, device=self.dev)
        ops.gemm_tc(a_hi=op[0], a_lo=op[1], a_dims=(B, 1, lk, inner), b_hi=aw["wout"][0], b_lo=aw["wout"][1], k=inner, n_cols=self.d, out_w=n,
                    z_outer=B, out_zo_stride=n * self.d, ldc=self.d, residual=residual, out_f32=out, bn=128, npass=self.npass)
        return out

    def embed(self, ids, cond_ids, batch):
        """-> (x fp32 [B*n_img, d], context planes [B*nc, d])."""
        B = cond_ids.shape[0]
        out = torch.empty((B, self.L, self.d), dtype=torch.float32, device=self.dev)
        a = EmbedArgs()
        self._keep = (ids.to(self.dev, torch.int64).contiguous(), cond_ids.to(self.dev, torch.int64).contiguous(),
                      batch["intrinsics_inv"].to(self.dev, torch.float32).contiguous(), batch["extrinsics_inv"].to(self.dev, torch.float32).contiguous())
        a.cam_idx, a.bev_idx = self._keep[0].data_ptr(), self._keep[1].data_ptr()
        a.intrinsics_inv, a.extrinsics_inv = self._keep[2].data_ptr(), self._keep[3].data_ptr()
        a.x_tok_emb, a.cond_tok_emb, a.x_pos_emb = self.tok_emb.data_ptr(), self.cond_tok_emb.data_ptr(), self.pos_emb.data_ptr()
        a.cond_static = self.cond_static.data_ptr()
        a.img_embed_w = self.img_w.data_ptr() if self.image_embed else None
        a.cam_embed_w = self.cam_w.data_ptr() if self.image_embed else None
        a.forward_shuffle_idx, a.pixel, a.out = self.order.data_ptr(), self.pixel.data_ptr(), out.data_ptr()
        a.B, a.ncam, a.hw, a.nc, a.n_img, a.L, a.d, a.vocab = B, self.cfg.num_cams, self.cfg.num_cam_tokens, self.nc, self.n_img, self.L, self.d, self.mask_id
        a.pad_last, a.bev_embed, a.row0, a.nrows = 0, int(self.bev_embed), 0, self.L
        ops.embed_assemble(a)
        x = out[:, self.nc:].reshape(B * self.n_img, self.d).contiguous()
        ctx = out[:, : self.nc].reshape(B * self.nc, self.d).contiguous()
        cp = self._planes((B * self.nc, self.d))
        ops.mg_head_planes(ctx, self.d, 0, self.nc, cp[0], cp[1], B, self.nc, self.d // 64)       # plain fp32 -> operand planes
        return x, cp

    @torch.no_grad()
    def forward(self, ids, cond_ids, batch):
        """ids [(b cam), hw] (mask id = vocab allowed), cond_ids [b, nc] -> (logits [(b cam), hw, vocab], embed [(b cam), hw, d])."""
        B, n, d, inner = cond_ids.shape[0], self.n_img, self.d, self.inner
        rows = B * n
        x, ctx = self.embed(ids, cond_ids, batch)
        for lw in self.layers:
            sa, ca, ff = lw["self"], lw["cross"], lw["ff"]
            yp = self._ln_planes(x, sa["gamma"], rows, f16f8=self.pair)
            qkv = torch.empty((rows, 3 * inner), dtype=torch.float32, device=self.dev)
            if self.pair:
                self._pair_linear(yp, sa["wqkv_p"], 3 * inner, rows, d, out_f32=qkv)
            else:
                self._linear(yp, sa["wqkv"], 3 * inner, rows, d, out_f32=qkv)
            if self.fused_self:
                x = self._attend_self_fused(qkv, sa, B, x)
            else:
                x = self._attend(qkv, 3 * inner, 0, qkv, 3 * inner, inner, 2 * inner, n, self.lk_self, sa, self.bias_self, self.mask_self, B, x)
            del qkv
            yp = self._ln_planes(x, ca["gamma"], rows, f16f8=self.pair)
            q = torch.empty((rows, inner), dtype=torch.float32, device=self.dev)
            if self.pair:
                self._pair_linear(yp, ca["wq_p"], inner, rows, d, out_f32=q)
            else:
                self._linear(yp, ca["wq"], inner, rows, d, out_f32=q)
            kv = torch.empty((B * self.nc, 2 * inner), dtype=torch.float32, device=self.dev)
            self._linear(ctx, ca["wkv"], 2 * inner, B * self.nc, d, out_f32=kv)
            if self.fused_self:
                x = self._attend_cross_fused(q, kv, ca, B, x)
            else:
                x = self._attend(q, inner, 0, kv, 2 * inner, 0, inner, self.nc, self.lk_cross, ca, self.bias_cross, self.mask_cross, B, x)
            yp = self._ln_planes(x, ff["gamma0"], rows, f16f8=self.pair)
            x2 = torch.empty_like(x)
            if self.pair:
                h = torch.empty((rows, self.n1_pad), dtype=torch.float32, device=self.dev)
                self._pair_linear(yp, ff["w1_p"], self.n1_pad, rows, d, out_f32=h)
                gp = (torch.empty((rows, self.f_pad), dtype=torch.float16, device=self.dev),
                      torch.empty((rows, 2 * self.f_pad), dtype=torch.uint8, device=self.dev))
                ops.mg_geglu_ln(h, ff["gamma3"], gp[0], gp[1], rows, self.f, self.f_pad, h_ld=self.n1_pad, f16f8=True)
                del h
                self._pair_linear(gp, ff["w2_p"], d, rows, self.f_pad, residual=x, out_f32=x2)
            else:
                h = torch.empty((rows, 2 * self.f), dtype=torch.float32, device=self.dev)
                self._linear(yp, ff["w1"], 2 * self.f, rows, d, out_f32=h)
                gp = self._planes((rows, self.f_pad))
                ops.mg_geglu_ln(h, ff["gamma3"], gp[0], gp[1], rows, self.f, self.f_pad)
                del h
                self._linear(gp, ff["w2"], d, rows, self.f_pad, residual=x, out_f32=x2)
            x = x2
        emb = torch.empty_like(x)
        ep = self._ln_planes(x, self.gamma_f, rows, y=emb)
        logits = torch.empty((rows, self.vocab), dtype=torch.float32, device=self.dev)
        self._linear(ep, self.w_logits, self.vocab, rows, d, out_f32=logits)
        self._last_embed_planes = ep
        hw = self.cfg.num_cam_tokens
        return logits.view(B * self.cfg.num_cams, hw, self.vocab), emb.view(B * self.cfg.num_cams, hw, d)

    @torch.no_grad()
    def critic_scores(self, ids, cond_ids, batch):
        """SelfCritic.forward_with_cond_scale (:377-379): Linear(d, 1) on the embeddings of a full forward -> [(b cam), hw]."""
        if self.critic is None:
            raise RuntimeError("no token critic weights were given")
        self.forward(ids, cond_ids, batch)
        rows = cond_ids.shape[0] * self.n_img
        out = torch.empty((rows, 1), dtype=torch.float32, device=self.dev)
        self._linear(self._last_embed_planes, self.critic[0], 1, rows, self.d, bias=self.critic[1], out_f32=out)
        return out.view(-1, self.cfg.num_cam_tokens)

    # ------------------------------------------------------------------ CUDA-graph replay of the two forwards of a generate step
    def _graphs(self, cond_ids, batch, use_critic):
        """One captured graph per (batch size, critic) holding [forward -> logits] and, with the critic, [forward -> scores] on static
        input buffers: a forward is ~400 launches, and at 2 scenes the launch overhead is a sixth of the step."""
        B = cond_ids.shape[0]
        key = (B, bool(use_critic))
        g = self._graph_cache.get(key) if hasattr(self, "_graph_cache") else None
        if g is None:
            if not hasattr(self, "_graph_cache"):
                self._graph_cache = {}
            st = {"ids": torch.full((B * self.cfg.num_cams, self.cfg.num_cam_tokens), self.mask_id, dtype=torch.int64, device=self.dev),
                  "cond": torch.zeros((B, self.nc), dtype=torch.int64, device=self.dev),
                  "batch": {"intrinsics_inv": torch.zeros((B, self.cfg.num_cams, 3, 3), device=self.dev),
                            "extrinsics_inv": torch.zeros((B, self.cfg.num_cams, 4, 4), device=self.dev)}}
            side = torch.cuda.Stream(self.dev)
            side.wait_stream(torch.cuda.current_stream(self.dev))
            with torch.cuda.stream(side):          # warm-up outside capture (lazy kernel attribute setup, allocator pools)
                self.forward(st["ids"], st["cond"], st["batch"])
                if use_critic:
                    self.critic_scores(st["ids"], st["cond"], st["batch"])
            torch.cuda.current_stream(self.dev).wait_stream(side)
            st["g_fwd"] = torch.cuda.CUDAGraph()
            with torch.cuda.graph(st["g_fwd"]):
                st["logits"], _ = self.forward(st["ids"], st["cond"], st["batch"])
            if use_critic:
                st["g_crit"] = torch.cuda.CUDAGraph()
                with torch.cuda.graph(st["g_crit"], pool=st["g_fwd"].pool()):
                    st["scores"] = self.critic_scores(st["ids"], st["cond"], st["batch"])
            self._graph_cache = {key: st}           # one live set of static buffers
            g = st
        g["cond"].copy_(cond_ids)
        g["batch"]["intrinsics_inv"].copy_(batch["intrinsics_inv"].to(self.dev, torch.float32))
        g["batch"]["extrinsics_inv"].copy_(batch["extrinsics_inv"].to(self.dev, torch.float32))
        return g

    @torch.no_grad()
    def generate(self, cond_ids, batch, timesteps=18, temperature=1.0, topk_filter_thres=0.9, critic_noise_scale=1.0, init_ids=None,
                 use_critic=None, noise=None, generator=None, trace=None, use_graph=True):
        """MaskGit.generate (:511-627).  noise(kind, step, shape) -> uniform(0, 1) tensor (tests replay the reference's draws); default:
        torch.rand on the device.  The token bookkeeping between the forwards runs in bevgen_mg_remask / bevgen_mg_sample."""
        dev = self.dev
        use_critic = (self.critic is not None) if use_critic is None else use_critic
        if noise is None:
            noise = lambda kind, step, shape: torch.rand(shape, device=dev, generator=generator)
        cond_ids = cond_ids.to(dev)
        ncam, hw = self.cfg.num_cams, self.cfg.num_cam_tokens
        shape = (cond_ids.shape[0] * ncam, hw)
        scores = torch.zeros(shape, dtype=torch.float32, device=dev)
        ids = torch.full(shape, self.mask_id, dtype=torch.long, device=dev)
        init_mask = None
        if init_ids is not None:
            init_ids = init_ids.to(dev)
            init_mask = init_ids != self.mask_id
        k = math.ceil((1 - topk_filter_thres) * self.vocab)
        gr = self._graphs(cond_ids, batch, use_critic) if use_graph else None
        # The cosine schedule is host (CPU tensor) arithmetic - no device read-back; the token bookkeeping between the forwards is two launches per
        # step: mg_remask (top-n_mask of the scores + critic noise -> mask_id, init_ids restored) and mg_sample (top-k filter, gumbel
        # arg-max, fill of the masked positions, and without a critic the next scores).
        ids = ids.contiguous()
        pending = None                                            # (uniform, scale) of the critic noise, folded into the next mg_remask
        init_c = None if init_ids is None else init_ids.contiguous()
        for step, (t, until_x0) in enumerate(zip(torch.linspace(0, 1, timesteps), reversed(range(timesteps)))):
            n_mask = max(int((torch.cos(t * math.pi * 0.5) * hw).item()), 1)      # CPU float32 arithmetic, exactly the reference's (:567); no device sync
            ops.mg_remask(scores.contiguous(), ids, n_mask, self.mask_id, uniform=None if pending is None else pending[0],
                          noise_scale=0.0 if pending is None else pending[1], init_ids=init_c)
            if gr is not None:
                gr["ids"].copy_(ids)
                gr["g_fwd"].replay()
                logits = gr["logits"]
            else:
                logits, _ = self.forward(ids, cond_ids, batch)
            if trace is not None:
                trace.append((ids.clone(), logits.clone()))
            temp = temperature * (until_x0 / timesteps)
            u = noise("gumbel", step, tuple(logits.shape)).to(dev, torch.float32).contiguous()
            nxt = None if use_critic else torch.empty(shape, dtype=torch.float32, device=dev)
            ops.mg_sample(logits.contiguous(), u, ids, k, 1.0 / max(temp, 1e-10), self.mask_id, scores=nxt)
            if use_critic:
                if gr is not None:
                    gr["ids"].copy_(ids)
                    gr["g_crit"].replay()
                    scores = gr["scores"]
                else:
                    scores = self.critic_scores(ids, cond_ids, batch)
                pending = (noise("critic", step, tuple(scores.shape)).to(dev, torch.float32).contiguous(), critic_noise_scale * (until_x0 / timesteps))
            else:
                scores, pending = nxt, None
        return ids.view(shape[0], self.cfg.cam_latent_h, self.cfg.cam_latent_w)
