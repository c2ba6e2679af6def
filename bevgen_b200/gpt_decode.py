"""KV-cache autoregressive sampler for the stage-2 transformer (the B200-native replacement of the reference's
O(L^2)-per-token loop, Net2NetTransformer.sample, modules/stage2/cond_transformer_multi_view.py:154-227).

Equivalence (checked by tests/test_decode_gpu.py against the reference's own outputs): with the closed-form
mask allowed(i,j) = (j < n_cond) or (i >= n_cond and j <= i), the 256 conditioning rows only see each other, so they are
prefilled once; every generated token then needs one new row attending to the cached keys with one camera-bias row.

Default path (batch <= 16, fp16 cache): ONE launch of the persistent kernel (csrc/decode_persistent.cu, bevgen_decode_persistent) runs all
remaining token steps - one CTA per SM, weights streamed once per step from a 3-byte fragment-ordered format, grid barriers between
phases, sampling and the next token's embedding inside the kernel.  The per-launch chain below is kept as the fallback for larger
batches / other cache types and as a cross-check in the tests (`sampler.persistent = False`).

Per step of the fallback (captured once as a CUDA graph and replayed 1535 times, the step counter lives in device memory):
  embed(row) -> [reduce+LayerNorm] -> 24 x { swap-AB tcgen05 GEMM (Wqkv) -> decode attention (append K/V, softmax, P.V, +residual)
  -> LayerNorm -> GEMM (W1) -> reduce+GELU -> GEMM (W2) -> reduce+bias+residual+LayerNorm } -> GEMM (head) -> top-k sample.
The weight GEMMs stream each weight matrix exactly once per step (HBM-bound: weights are the 128-row MMA operand, the 16
batch rows are the N=16 operand); split-K spreads every matrix over ~100 CTAs.
"""
import ctypes as C
import math

import torch

from . import _lib, ops
from .ops import EmbedArgs, _ptr, _stream


def _kslice(K):
    for k in (256, 128, 64):
        if K % k == 0:
            return k
    raise ValueError(f"reduction length {K} is not a multiple of 64")


class GPTSampler:
    def __init__(self, engine, batch_size, kv_dtype=None):
        self.eng = e = engine
        if not e.decode_causal:
            raise NotImplementedError("KV-cache decoding needs the [cond | causal] mask on the real tokens (causal_order=True)")
        self.B = B = batch_size
        self.Bp = ((B + 15) // 16) * 16
        dev, d, H = e.dev, e.d, e.H
        # parity mode: fp16 cache (3.4e-5 max logit error over all 1536 replayed steps vs 1.7e-5 for an fp32 cache, half the bytes)
        kvt = kv_dtype or (torch.float16 if e.npass == 3 else torch.bfloat16)
        self.kv_bf16 = {torch.float32: 0, torch.bfloat16: 1, torch.float16: 2}[kvt]      # C ABI cache-type code
        nl = len(e.layers)
        self.Lmax = ((e.L + 127) // 128) * 128          # K cache is blocked by 128 keys
        self.kc = [torch.zeros((B, H, self.Lmax // 128, 64, 128), dtype=kvt, device=dev) for _ in range(nl)]
        self.vc = [torch.zeros((B, H, self.Lmax, 64), dtype=kvt, device=dev) for _ in range(nl)]
        self.step = torch.zeros(1, dtype=torch.int32, device=dev)
        self.cam_idx = torch.full((B, e.cfg.num_cams, e.cfg.num_cam_tokens), e.cfg.vocab_size, dtype=torch.int64, device=dev)
        self.tokens = torch.zeros((B, e.n_img), dtype=torch.int64, device=dev)
        f32 = lambda *s: torch.zeros(s, dtype=torch.float32, device=dev)
        self.xrow, self.y, self.x1 = f32(B, 1, d), f32(self.Bp, d), f32(self.Bp, d)
        pl = lambda n: (torch.zeros((self.Bp, n), dtype=torch.bfloat16, device=dev),
                        torch.zeros((self.Bp, n), dtype=torch.bfloat16, device=dev) if e.npass == 3 else None)
        self.yp, self.zp, self.hp, self.fp = pl(d), pl(d), pl(4 * d), pl(d)
        self.vpad = e.whead[0].shape[0]
        ks = lambda K: K // _kslice(K)
        self.ks_d, self.ks_4d = ks(d), ks(4 * d)
        self.part_qkv = f32(self.ks_d, self.Bp, 3 * d)
        self.part_h = f32(self.ks_d, self.Bp, 4 * d)
        self.part_o = f32(self.ks_4d, self.Bp, d)
        self.part_v = f32(self.ks_d, self.Bp, self.vpad)
        self.bias_cc = None if e.bias is None else e.bias[: e.nc, : e.nc].contiguous()
        self.mask_cc = e.mask_u8[: e.nc, : e.nc].contiguous()
        full_cc = bool(self.mask_cc.all())      # cond rows see every cond column: the fused kernel's "all cond" case
        self.bias_cc_f16 = ops.tile_attention_bias(torch.zeros((e.nc, e.nc), device=dev) if self.bias_cc is None else self.bias_cc,
                                                   float(e.dh) ** -0.5) if (full_cc and e.nc % 128 == 0) else None
        # per-layer block layouts (density < 1): full [H][nb][nb] for the decode rows, contiguous cond x cond corner for the prefill
        ncb = e.nc // e.layout_block
        self.lay_cc = None if e.layouts is None else [l[:, :ncb, :ncb].contiguous() for l in e.layouts]
        self.attn_ws = f32(_lib.load().bevgen_dec_attention_workspace_floats(B, H))
        self.attn_cnt = torch.zeros(B * H, dtype=torch.int32, device=dev)
        self.row_cnt = torch.zeros(B, dtype=torch.int32, device=dev)
        self.cnt_h = torch.zeros(4 * d // 128 + 1, dtype=torch.int32, device=dev)      # MLP1 finalize tickets
        self.cnt_o = torch.zeros(d // 128 + 1, dtype=torch.int32, device=dev)          # MLP2 finalize tickets
        self.x2 = f32(self.Bp, d)
        self.fuse_finalize = False       # split-K finalize inside the GEMMs (one tail CTA) measured slower than separate reduce kernels
        self.fuse_ln2 = True             # ln2 applied by the last head CTA of the decode-attention kernel
        self.use_pdl = True              # programmatic dependent launch along the decode chain (bevgen_set_pdl)
        self.persistent = bool(B <= 16 and kvt == torch.float16)      # one persistent launch for all token steps
        self.profile_phases = False      # per-phase %globaltimer marks inside the persistent kernel (each read costs ~0.3 us: off by default)
        self._pk = None
        self.graph = None
        self._graph_key = None
        self._graph_launches = 0
        self.trace = None
        self.timing = None               # bench hook: a list -> sample() appends a (start, end) CUDA-event pair around its decode loop
        # bench / roofline metadata of the decode loop
        self.ncu_traffic_bytes = None
        self.ncu_traffic_source = None

    PROFILE_SLOTS = ("qkv", "qkv_barrier", "attention", "attention_barrier", "mlp1", "mlp1_barrier", "mlp2", "mlp2_barrier", "head", "head_barrier",
                     "sample_embed", "sample_barrier",
                     "fine_afrag", "fine_linear_ring_wait", "fine_linear_math", "fine_attn_prologue", "fine_attn_wait_full_w0", "fine_attn_units_w0", "fine_attn_merge",
                     "fine_mlp2", "fine_attn_wait_prev_w0", "fine_attn_end_barrier", "producer_ring_full_weights", "producer_ring_full_attn", "producer_epoch_gate")

    def last_profile(self):
        """Per-phase milliseconds of the last persistent launch (mean over CTAs; the kernel's own %globaltimer marks)."""
        if not self._pk or "prof" not in self._pk:
            return None
        t = self._pk["prof"][: -32].double().mean(0) / 1e6
        return {k: float(v) for k, v in zip(self.PROFILE_SLOTS, t.tolist())}

    def last_failure(self):
        """(code, cta, step, layer, phase, a, b, thread) left by a timed-out wait of the persistent kernel, or None."""
        if not self._pk or "dbg" not in self._pk or int(self._pk["dbg"][0]) == 0:
            return None
        return tuple(int(v) & 0xffffffff for v in self._pk["dbg"].tolist())

    @property
    def kernel_name(self):
        return ("decode_persistent_kernel (one launch for all token steps)" if self.persistent else
                "decode step = CUDA graph of swap-AB tcgen05 GEMMs + dec_attn_kernel + dec_reduce_* (146 launches per token)")

    @property
    def launches_per_token(self):
        return 1.0 / max(self.eng.n_img - 1, 1) if self.persistent else 146

    # ------------------------------------------------------------------ persistent kernel
    def _pack_linear(self, w, n_quarters=1):
        """fp32 [rows][ld] -> the kernel's 3-byte format (fp16 plane + e4m3 plane of (w - fp16(w)) * 2^e), bevgen_pack_decode_linear."""
        lib, e = _lib.init(), self.eng
        w = w.detach().to(e.dev, torch.float32).contiguous()
        rows, ld = w.shape
        amax = float(w.abs().max())
        ex = 0 if amax == 0.0 else min(max(19 - math.floor(math.log2(amax)), -20), 40)      # max |residual| * 2^ex <= 256 (e4m3 saturates at 448)
        lo_mul = 2.0 ** ex
        nbytes = lib.bevgen_pack_decode_linear(None, rows, ld, e.d, n_quarters, lo_mul, None, None)
        if nbytes <= 0:
            _lib.check(int(nbytes), "pack_decode_linear")
        out = torch.empty(int(nbytes), dtype=torch.uint8, device=e.dev)
        ops.Stats.launches += 1
        rc = lib.bevgen_pack_decode_linear(_ptr(w), rows, ld, e.d, n_quarters, lo_mul, _ptr(out), _stream())
        if rc != nbytes:
            _lib.check(int(rc), "pack_decode_linear")
        return out, 1.0 / lo_mul

    def _fold_ln(self, w, gamma, beta, bias):
        """Lazy LayerNorm (decode_persistent.cu): LN(x) W^T + b = rstd * (x W'^T - mean * c1) + c2 with W' = W * gamma (per input column),
        c1_n = sum_k W'_nk, c2_n = b_n + sum_k beta_k W_nk.  The constants are summed in fp64."""
        e = self.eng
        w = w.detach().to(e.dev, torch.float32)
        wg = (w * gamma.to(e.dev, torch.float32)[None, :]).contiguous()
        c1 = wg.double().sum(1).float()
        c2 = (w.double() @ beta.to(e.dev).double()).float()
        if bias is not None:
            c2 = c2 + bias.detach().to(e.dev, torch.float32)
        rows = ((w.shape[0] + 7) // 8) * 8
        if rows != w.shape[0]:
            c1 = torch.cat([c1, c1.new_zeros(rows - w.shape[0])])
            c2 = torch.cat([c2, c2.new_zeros(rows - w.shape[0])])
        return wg, c1.contiguous(), c2.contiguous()

    def _build_persistent(self):
        lib, e = _lib.init(), self.eng
        sd, d = e.sd, e.d
        keep, arr = [], (_lib.DecodeLayer * len(e.layers))()
        for i, lw in enumerate(e.layers):
            p = f"blocks.{i}"
            wqkv = torch.cat([sd[f"{p}.attention.{n}.weight"].detach().to(e.dev, torch.float32) for n in ("query", "key", "value")], 0)
            wq, c1q, c2q = self._fold_ln(wqkv, lw["ln1"][0], lw["ln1"][1], lw["bqkv"])
            pq, sq = self._pack_linear(wq)
            del wqkv, wq
            w1, c11, c21 = self._fold_ln(sd[f"{p}.mlp.0.weight"], lw["ln2"][0], lw["ln2"][1], lw["b1"])
            p1, s1 = self._pack_linear(w1)
            del w1
            p2, s2 = self._pack_linear(sd[f"{p}.mlp.2.weight"], n_quarters=4)
            keep += [pq, p1, p2, c1q, c2q, c11, c21]
            a = arr[i]
            a.w_qkv, a.w_1, a.w_2 = pq.data_ptr(), p1.data_ptr(), p2.data_ptr()
            a.c1_qkv, a.c2_qkv, a.c2_2 = c1q.data_ptr(), c2q.data_ptr(), lw["b2"].data_ptr()
            a.ln1_g, a.ln1_b, a.c1_1, a.c2_1 = lw["ln1"][0].data_ptr(), lw["ln1"][1].data_ptr(), c11.data_ptr(), c21.data_ptr()
            a.k_cache, a.v_cache = self.kc[i].data_ptr(), self.vc[i].data_ptr()
            a.layout = None if lw.get("layout") is None else lw["layout"].data_ptr()
            a.s_qkv, a.s_1, a.s_2 = sq, s1, s2
        wh, c1h, c2h = self._fold_ln(sd["head.weight"], e.ln_f[0], e.ln_f[1], None)
        ph, sh = self._pack_linear(wh)
        del wh
        layers_dev = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8).to(e.dev)
        nf, ncnt = C.c_longlong(), C.c_longlong()
        _lib.check(lib.bevgen_decode_workspace(self.B, d, e.H, e.vocab, C.byref(nf), C.byref(ncnt)), "decode_workspace")
        ws = torch.zeros(nf.value, dtype=torch.float32, device=e.dev)
        cnt = torch.zeros(ncnt.value, dtype=torch.int32, device=e.dev)
        self._pk = dict(keep=keep, layers=layers_dev, head=ph, s_head=sh, c1_head=c1h, c2_head=c2h, ws=ws, cnt=cnt,
                        weight_bytes=sum(t.numel() for t in keep[0::7] + keep[1::7] + keep[2::7]) + ph.numel())

    def _run_persistent(self, batch, step_begin, step_end, temperature, top_k, greedy, seed, forced):
        lib, e = _lib.init(), self.eng
        if self._pk is None:
            self._build_persistent()
        pk = self._pk
        a = _lib.DecodeArgs()
        a.layers, a.n_layers = pk["layers"].data_ptr(), len(e.layers)
        a.w_head, a.s_head, a.c1_head, a.c2_head = pk["head"].data_ptr(), pk["s_head"], pk["c1_head"].data_ptr(), pk["c2_head"].data_ptr()
        a.batch, a.d, a.heads, a.vocab, a.n_cond, a.n_img, a.lmax = self.B, e.d, e.H, e.vocab, e.nc, e.n_img, self.Lmax
        a.ncam, a.hw, a.step_begin, a.step_end = e.cfg.num_cams, e.cfg.num_cam_tokens, step_begin, step_end
        a.cam_idx, a.x_tok_emb, a.x_pos_emb = self.cam_idx.data_ptr(), e.x_tok_emb.data_ptr(), e.x_pos_emb.data_ptr()
        if e.image_embed:
            a.img_embed_w, a.cam_embed_w = e.img_w.data_ptr(), e.cam_w.data_ptr()
            a.intrinsics_inv, a.extrinsics_inv, a.pixel = batch["intrinsics_inv"].data_ptr(), batch["extrinsics_inv"].data_ptr(), e.pixel.data_ptr()
        a.forward_shuffle_idx = e.fwd.data_ptr()
        a.camera_bias, a.bias_ld = (None, 0) if e.bias is None else (e.bias.data_ptr(), e.L)
        a.scale, a.temperature, a.top_k, a.greedy, a.seed = float(e.dh) ** -0.5, float(temperature), int(top_k or 0), int(greedy), int(seed) & (2 ** 64 - 1)
        a.forced_tokens = None if forced is None else forced.data_ptr()
        a.tokens_out = self.tokens.data_ptr()
        a.logits_trace = None if self.trace is None else self.trace.data_ptr()
        a.layout_block = e.layout_block
        a.layout_ld = 0 if e.layouts is None else e.layouts.shape[-1]
        a.workspace, a.counters = pk["ws"].data_ptr(), pk["cnt"].data_ptr()
        if "dbg" not in pk:
            pk["dbg"] = torch.zeros(8, dtype=torch.int32).pin_memory()           # readable by the host after a time-out trap
            pk["prof"] = torch.zeros((lib.bevgen_sm_count() + 32, 32), dtype=torch.int64, device=e.dev)      # + 1024 trace entries (BEVGEN_DP_DBG & 64)
        pk["dbg"].zero_()
        a.debug, a.profile = pk["dbg"].data_ptr(), (pk["prof"].data_ptr() if self.profile_phases else None)
        ops.Stats.launches += 1
        _lib.check(lib.bevgen_decode_persistent(C.byref(a), _stream()), "decode_persistent")

    # ------------------------------------------------------------------ launches
    def _gemm_t(self, w, xp, n_out, K, part, fin=None):
        """partials[z][b][n] = sum_{k in slice z} W[n][k] x[b][k]  — weights are the 128-row (M) operand, batch the N=16 operand."""
        kk = _kslice(K)
        ops.gemm_tc(a_hi=w[0], a_lo=w[1], a_dims=(1, 1, w[0].shape[0], K), b_hi=xp[0], b_lo=xp[1], k=kk, n_cols=self.B, a_c_zstride=kk,
                    b_k_zstride=kk, z_inner=K // kk, out_w=n_out, out_zi_stride=self.Bp * n_out, ldc=n_out, out_f32=part,
                    flags=ops.GF_OUT_T, bn=16, npass=self.eng.npass, algo_flops=2.0 * self.B * n_out * K, fin=fin)

    def _reduce_ln(self, part, ks, bias, residual, res_stride, ln, y, planes, rows=None, x_out=None):
        lib = _lib.init()
        ops.Stats.launches += 1
        d = self.eng.d
        zs = 0 if part is None else part.shape[1] * part.shape[2]
        _lib.check(lib.bevgen_dec_reduce_ln(_ptr(part), ks, zs, _ptr(bias), _ptr(residual), res_stride, _ptr(ln[0]), _ptr(ln[1]), 1e-5, _ptr(x_out),
                                            _ptr(y), _ptr(planes[0]) if planes else None, _ptr(planes[1]) if planes else None,
                                            self.B if rows is None else rows, d, _stream()), "dec_reduce_ln")

    def _sample(self, temperature, top_k, greedy, seed, forced):
        lib, e = _lib.init(), self.eng
        ops.Stats.launches += 2
        _lib.check(lib.bevgen_sample_topk(_ptr(self.part_v), self.ks_d, self.Bp * self.vpad, self.vpad, e.vocab, float(temperature),
                                          int(top_k or 0), int(greedy), C.c_ulonglong(seed), _ptr(forced), _ptr(e.fwd), _ptr(self.cam_idx),
                                          _ptr(self.tokens), _ptr(self.trace), None, _ptr(self.step), self.B, e.n_img, e.cfg.num_cam_tokens,
                                          e.cfg.num_cams, _stream()), "sample_topk")
        _lib.check(lib.bevgen_dec_advance(_ptr(self.step), _stream()), "dec_advance")

    def _embed_args(self, bev_idx, batch):
        e = self.eng
        a = EmbedArgs()
        self._keep = [bev_idx, batch["intrinsics_inv"], batch["extrinsics_inv"]]
        a.cam_idx, a.bev_idx = self.cam_idx.data_ptr(), bev_idx.data_ptr()
        a.intrinsics_inv, a.extrinsics_inv = batch["intrinsics_inv"].data_ptr(), batch["extrinsics_inv"].data_ptr()
        a.x_tok_emb, a.cond_tok_emb = e.x_tok_emb.data_ptr(), e.cond_tok_emb.data_ptr()
        a.x_pos_emb, a.cond_static = e.x_pos_emb.data_ptr(), e.cond_static.data_ptr()
        a.img_embed_w = e.img_w.data_ptr() if e.image_embed else None
        a.cam_embed_w = e.cam_w.data_ptr() if e.image_embed else None
        a.forward_shuffle_idx, a.pixel, a.out = e.fwd.data_ptr(), e.pixel.data_ptr(), self.xrow.data_ptr()
        a.step_ptr = self.step.data_ptr()
        a.B, a.ncam, a.hw, a.nc, a.n_img, a.L, a.d, a.vocab = self.B, e.cfg.num_cams, e.cfg.num_cam_tokens, e.nc, e.n_img, e.L, e.d, e.cfg.vocab_size
        a.pad_last, a.bev_embed, a.row0, a.nrows = 0, int(e.bev_embed), 0, 1
        return a

    def _step(self, embed_args, temperature, top_k, greedy, seed, forced):
        """One decode step (graph-capturable: no allocation, no host sync, step counter read on the device)."""
        lib = _lib.init()
        old = lib.bevgen_set_pdl(int(self.use_pdl))
        try:
            self._step_body(embed_args, temperature, top_k, greedy, seed, forced)
        finally:
            lib.bevgen_set_pdl(old)

    def _step_body(self, embed_args, temperature, top_k, greedy, seed, forced):
        e, lib = self.eng, _lib.init()
        d, H = e.d, e.H
        ops.embed_assemble(embed_args)
        self._reduce_ln(None, 0, None, self.xrow, d, e.layers[0]["ln1"], self.y, self.yp)
        for li, lw in enumerate(e.layers):
            self._gemm_t(lw["wqkv"], self.yp, 3 * d, d, self.part_qkv)
            ops.Stats.launches += 1
            fz = self.fuse_finalize
            f2 = self.fuse_ln2 or fz
            _lib.check(lib.bevgen_dec_attention(_ptr(self.part_qkv), self.ks_d, self.Bp * 3 * d, _ptr(lw["bqkv"]), _ptr(self.y), _ptr(e.bias), e.L,
                                                _ptr(self.kc[li]), _ptr(self.vc[li]), self.kv_bf16, _ptr(self.x1), _ptr(self.step), _ptr(self.attn_ws),
                                                _ptr(self.attn_cnt), self.B, e.nc, H, d, self.Lmax, float(e.dh) ** -0.5,
                                                _ptr(self.row_cnt) if f2 else None, _ptr(lw["ln2"][0]) if f2 else None,
                                                _ptr(lw["ln2"][1]) if f2 else None, 1e-5, _ptr(self.zp[0]) if f2 else None,
                                                _ptr(self.zp[1]) if f2 else None, _ptr(lw.get("layout")), e.layout_block,
                                                0 if lw.get("layout") is None else lw["layout"].shape[-1], _stream()), "dec_attention")
            last = li == len(e.layers) - 1
            nxt = e.ln_f if last else e.layers[li + 1]["ln1"]
            if fz:
                # split-K reductions fused into the producing GEMMs (last CTA per feature tile): 4 launches per layer
                self._gemm_t(lw["w1"], self.zp, 4 * d, d, self.part_h,
                             fin=dict(mode=1, gelu=1, rows=self.B, counters=self.cnt_h, hi=self.hp[0], lo=self.hp[1], bias=lw["b1"]))
                outp = self.fp if last else self.yp
                self._gemm_t(lw["w2"], self.hp, d, 4 * d, self.part_o,
                             fin=dict(mode=2, rows=self.B, counters=self.cnt_o, hi=outp[0], lo=outp[1], bias=lw["b2"], resid=self.x1, x=self.x2,
                                      y=None if last else self.y, gamma=nxt[0], beta=nxt[1], eps=1e-5))
                continue
            if not f2:
                self._reduce_ln(None, 0, None, self.x1, d, lw["ln2"], None, self.zp)
            self._gemm_t(lw["w1"], self.zp, 4 * d, d, self.part_h)
            ops.Stats.launches += 1
            _lib.check(lib.bevgen_dec_reduce_act(_ptr(self.part_h), self.ks_d, self.Bp * 4 * d, _ptr(lw["b1"]), 1, _ptr(self.hp[0]), _ptr(self.hp[1]),
                                                 self.B, 4 * d, _stream()), "dec_reduce_act")
            self._gemm_t(lw["w2"], self.hp, d, 4 * d, self.part_o)
            self._reduce_ln(self.part_o, self.ks_4d, lw["b2"], self.x1, d, nxt, None if last else self.y, self.fp if last else self.yp)
        self._gemm_t(e.whead, self.fp, self.vpad, d, self.part_v)
        self._sample(temperature, top_k, greedy, seed, forced)

    # ------------------------------------------------------------------ prefill
    def _prefill(self, bev_idx, batch):
        e = self.eng
        B, nc, d = self.B, e.nc, e.d
        x = e.embed(self.cam_idx, bev_idx, batch, sampling=True, row0=0, nrows=nc)
        lib = _lib.init()
        for li, lw in enumerate(e.layers):
            def store(qkv, li=li):
                ops.Stats.launches += 1
                _lib.check(lib.bevgen_kv_store(_ptr(qkv[0]), _ptr(qkv[1]), _ptr(self.kc[li]), _ptr(self.vc[li]), self.kv_bf16, B, nc, nc, e.H, d,
                                               self.Lmax, _stream()), "kv_store")
            x = e.block(x, lw, B, nc, attn_kw=dict(bias=self.bias_cc, mask=self.mask_cc, causal=False, allowed=float(nc * nc),
                                                    fused_cond=self.bias_cc_f16, layout=None if self.lay_cc is None else self.lay_cc[li]),
                        on_qkv=store)
        # logits of decode-order token 0 come from the last conditioning row (mingpt_sparse.py:390)
        self._last = x
        self._reduce_ln(None, 0, None, x.view(-1)[(nc - 1) * d:], nc * d, e.ln_f, None, self.fp)
        self._gemm_t(e.whead, self.fp, self.vpad, d, self.part_v)

    # ------------------------------------------------------------------ public
    @torch.no_grad()
    def sample(self, bev_idx, batch, temperature=1.0, top_k=None, greedy=False, seed=0, forced_tokens=None, steps=None,
               trace_logits=False, use_graph=True):
        """-> tokens int64 (B, num_cams, cam_tokens) [, logits trace (steps, B, vocab)].  forced_tokens (B, n_img) in decode
        order turns this into a teacher-forced replay (the sampled token is replaced, logits are still traced)."""
        e = self.eng
        dev = e.dev
        steps = e.n_img if steps is None else steps
        bev_idx = bev_idx.to(dev, torch.int64).contiguous()
        batch = {k: batch[k].to(dev, torch.float32).contiguous() for k in ("intrinsics_inv", "extrinsics_inv")}
        forced = None if forced_tokens is None else forced_tokens.to(dev, torch.int64).contiguous()
        self.trace = torch.zeros((e.n_img, self.B, e.vocab), dtype=torch.float32, device=dev) if trace_logits else None
        self.cam_idx.fill_(e.cfg.vocab_size)
        self.step.zero_()
        self._prefill(bev_idx, batch)
        self._sample(temperature, top_k, greedy, seed, forced)           # token 0; step -> 1
        args = self._embed_args(bev_idx, batch)
        key = (bev_idx.data_ptr(), batch["intrinsics_inv"].data_ptr(), batch["extrinsics_inv"].data_ptr(), float(temperature), top_k, greedy,
               seed, None if forced is None else forced.data_ptr(), None if self.trace is None else self.trace.data_ptr())
        done = 1
        if self.timing is not None:
            ev0 = torch.cuda.Event(enable_timing=True)
            ev0.record()
        if self.persistent and steps > 1:
            self._hold = (bev_idx, batch, forced, self.trace)
            self._run_persistent(batch, 1, steps, temperature, top_k, greedy, seed, forced)
            done = steps
        if steps > done and done == 1:
            self._step(args, temperature, top_k, greedy, seed, forced)    # eager step 1 (also warms every kernel variant up)
            done = 2
        if use_graph and steps > done:
            if self.graph is None or self._graph_key != key:
                self._hold = (bev_idx, batch, forced, self.trace)         # keep the captured buffers alive
                g = torch.cuda.CUDAGraph()
                snap = (self.step.clone(), self.cam_idx.clone(), self.tokens.clone())
                n0 = ops.Stats.launches
                with torch.cuda.graph(g):
                    self._step(args, temperature, top_k, greedy, seed, forced)
                self._graph_launches = ops.Stats.launches - n0
                ops.Stats.launches = n0                                   # capture does not launch
                # capture does not execute, but be safe if a driver replays it: restore the state
                self.step.copy_(snap[0]); self.cam_idx.copy_(snap[1]); self.tokens.copy_(snap[2])
                self.graph, self._graph_key = g, key
            for _ in range(done, steps):
                self.graph.replay()
            ops.Stats.launches += (steps - done) * self._graph_launches
        else:
            for _ in range(done, steps):
                self._step(args, temperature, top_k, greedy, seed, forced)
        if self.timing is not None:
            ev1 = torch.cuda.Event(enable_timing=True)
            ev1.record()
            self.timing.append((ev0, ev1))
        out = self.cam_idx.clone()
        return (out, self.trace[:steps]) if trace_logits else out

    def bytes_per_batch(self, steps=None):
        """Algorithmic HBM bytes of one full sample(): weights streamed once per step + KV cache reads."""
        e = self.eng
        steps = e.n_img if steps is None else steps
        wbytes = (3 if self.persistent else (2 if e.npass == 1 else 4)) * (len(e.layers) * 12 * e.d * e.d + e.vocab * e.d)
        kvb = 2 if self.kv_bf16 else 4        # bf16 / fp16 caches are 2 bytes per element
        kv = sum(len(e.layers) * 2 * (e.nc + t) * e.d * kvb for t in range(1, steps)) * self.B
        return wbytes * steps + kv
