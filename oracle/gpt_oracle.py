"""TEST INFRASTRUCTURE — CPU fp32 restatement of the reference's stage-2 transformer inference path.

Functional (state-dict driven) restatement of, under /root/reference/multi_view_generation:
  GPT.forward                    modules/transformer/mingpt_sparse.py:319-391
  Block.forward                  mingpt_sparse.py:240-253  (residual taken from ln1(x), NOT x)
  CustomSparseSelfAttention      mingpt_sparse.py:185-212  (q/k/v Linear, 16 heads, NO output projection)
  SparseSelfAttention.forward    modules/transformer/sparse_self_attention.py:128-177, dense fp32:
       P = softmax( d_head^-1/2 * (Q K^T + bias) ) over (layout block present) & (attn_mask != 0);  O = P V
       (DeepSpeed 0.7.4 Triton ops are absent offline: pinned_requirements.txt:5)
  Net2NetTransformer.sample tail modules/stage2/cond_transformer_multi_view.py:138-142,200-219
plus a KV-cache formulation of `sample` used to check the decode kernels.
`geo` carries the host-side artefacts (attention_mask, prob_matrix, forward/backward_shuffle_idx, sizes)
— in tests these come from bevgen_b200.GPTConfig, itself pinned against tests/golden/gptconfig_*.pt.
Pinned against the reference by tests/golden/gpt_*.pt. Not shipped, never on the product path.
"""
import math

import torch
import torch.nn.functional as F


def generate_grid(height, width):
    """mingpt_sparse.py:256-264 -> (1,3,h,w): x in [0,1] along w, y along h, ones."""
    xs = torch.linspace(0, 1, width)
    ys = torch.linspace(0, 1, height)
    g = torch.stack(torch.meshgrid((xs, ys), indexing="xy"), 0)
    g = F.pad(g, (0, 0, 0, 0, 0, 1), value=1)
    return g[None]


def bev_grid(bev_h, bev_w, offset=0):
    """mingpt_sparse.get_bev_grid :116-141 -> (3,h,w) ego-frame metres."""
    grid = generate_grid(bev_h, bev_w).squeeze(0)
    grid[0] = bev_w * grid[0]
    grid[1] = bev_h * grid[1]
    sh, sw = bev_h / 80, bev_w / 80
    V = torch.tensor([[0., -sw, bev_w / 2.], [-sh, 0., bev_h * offset + bev_h / 2.], [0., 0., 1.]])
    out = V.inverse() @ grid.reshape(3, -1)
    return out.reshape(3, bev_h, bev_w)


def camera_bias(sd, geo):
    """mingpt_sparse.py:375-380: scatter the tril parameter vector, add the prior (float64 -> activation dtype)."""
    L = geo["gpt_block_size"]
    idx = torch.tril_indices(L, L)
    b = torch.zeros(1, L, L, dtype=torch.float32)
    b[:, idx[0], idx[1]] = sd["camera_bias_emb"]
    return b + geo["prob_matrix"].to(torch.float32)


def embed(sd, geo, cam_indices, bev_indices, batch, sampling):
    """mingpt_sparse.py:319-373 -> (B, L, d) input embeddings in sequence order [cond | img(decode order) | pad]."""
    I_inv, E_inv = batch["intrinsics_inv"].float(), batch["extrinsics_inv"].float()
    b, ncam = I_inv.shape[:2]
    h, w = geo["cam_latent_res"]
    d = sd["x_tok_emb.weight"].shape[1]
    cam_indices = cam_indices.clone()
    if not sampling:
        cam_indices[:, -1, -1] = geo["vocab_size"]
    x = sd["x_tok_emb.weight"][cam_indices]                       # b cam hw d
    c_embed = None
    if "img_embed.weight" in sd:
        plane = generate_grid(h, w)[None]                           # 1 1 3 h w
        plane[:, :, 0] *= geo["cam_res"][0]
        plane[:, :, 1] *= geo["cam_res"][1]
        c = E_inv[..., -1:]                                         # b cam 4 1
        c_embed = torch.einsum("dk,bnk->bnd", sd["cam_embed.weight"].reshape(d, 4), c[..., 0])   # b cam d
        pix = plane.reshape(1, 1, 3, h * w)
        cam = I_inv @ pix                                           # b cam 3 hw
        cam = F.pad(cam, (0, 0, 0, 1), value=1)
        ray = E_inv @ cam                                           # b cam 4 hw
        d_embed = torch.einsum("dk,bnkp->bnpd", sd["img_embed.weight"].reshape(d, 4), ray)       # b cam hw d
        e = d_embed - c_embed[:, :, None, :]
        e = e / (e.norm(dim=-1, keepdim=True) + 1e-7)
        x = x + e
    cond = sd["cond_tok_emb.weight"][bev_indices]                 # b nc d
    if "bev_embed.weight" in sd:
        bh, bw = geo["bev_latent_res"]
        g = bev_grid(bh, bw)[:2].reshape(2, bh * bw)                # 2 nc
        grid_embed = (sd["bev_embed.weight"].reshape(d, 2) @ g).t() + sd["bev_embed.bias"]       # nc d
        bev_cam = (sd["bev_cam_pos_emb"] + c_embed[:, :, None, :]).sum(1)                         # b nc d
        cond = cond + (grid_embed[None] - bev_cam)
    x = x.reshape(b, ncam * h * w, d) + sd["x_pos_emb"][:, : ncam * h * w]
    cond = cond + sd["cond_pos_emb"]
    x = x[:, geo["forward_shuffle_idx"]]
    seq = torch.cat([cond, x], 1)
    L = geo["gpt_block_size"]
    if seq.shape[1] < L:
        pad = sd["x_tok_emb.weight"][geo["vocab_size"]].expand(b, L - seq.shape[1], d)
        seq = torch.cat([seq, pad], 1)
    return seq


def attention(q, k, v, bias, mask, layout=None, block=16):
    """Dense restatement of sparse_self_attention.py:153-176. q,k,v (B,H,L,dh); bias (1,L,L) or None; mask (L,L) {0,1}."""
    dh = q.shape[-1]
    s = q @ k.transpose(-1, -2)
    if bias is not None:
        s = s + bias[:, None]
    s = s * (float(dh) ** -0.5)
    if layout is not None:
        dense = layout.bool().repeat_interleave(block, -2).repeat_interleave(block, -1)
        s = s.masked_fill(~dense[None], float("-inf"))
    s = s.masked_fill(mask[None, None] == 0, float("-inf"))
    return torch.softmax(s, -1) @ v


def block(x, sd, i, nh, bias, mask, layout=None, blk=16):
    p = f"blocks.{i}"
    d = x.shape[-1]
    x = F.layer_norm(x, (d,), sd[f"{p}.ln1.weight"], sd[f"{p}.ln1.bias"])          # overwrites the stream (:242)
    B, L, _ = x.shape

    def proj(n):
        return F.linear(x, sd[f"{p}.attention.{n}.weight"], sd[f"{p}.attention.{n}.bias"]).view(B, L, nh, d // nh).permute(0, 2, 1, 3)

    a = attention(proj("query"), proj("key"), proj("value"), bias, mask, layout, blk)
    x = x + a.permute(0, 2, 1, 3).reshape(B, L, d)
    y = F.layer_norm(x, (d,), sd[f"{p}.ln2.weight"], sd[f"{p}.ln2.bias"])
    y = F.linear(F.gelu(F.linear(y, sd[f"{p}.mlp.0.weight"], sd[f"{p}.mlp.0.bias"])), sd[f"{p}.mlp.2.weight"], sd[f"{p}.mlp.2.bias"])
    return x + y


def forward(sd, geo, cam_indices, bev_indices, batch, sampling, layouts=None, return_hidden=False):
    """GPT.forward -> logits (B, num_img_tokens, vocab) in (cam,h,w) order."""
    x = embed(sd, geo, cam_indices, bev_indices, batch, sampling)
    bias = camera_bias(sd, geo) if "camera_bias_emb" in sd else None
    mask = geo["attention_mask"]
    hidden = []
    for i in range(geo["num_layers"]):
        x = block(x, sd, i, geo["num_heads"], bias, mask, None if layouts is None else layouts[i], geo["sparse_block_size"])
        if return_hidden:
            hidden.append(x)
    d = x.shape[-1]
    x = F.layer_norm(x, (d,), sd["ln_f.weight"], sd["ln_f.bias"])
    logits = F.linear(x, sd["head.weight"])
    pad = geo["num_pad_tokens"]
    if pad:
        logits = logits[:, :-pad]
    nc = geo["num_cond_tokens"]
    ret = logits[:, nc - 1:-1][:, geo["backward_shuffle_idx"]]
    return (ret, hidden) if return_hidden else ret


def top_k_logits(logits, k):
    """cond_transformer_multi_view.py:138-142: everything strictly below the k-th value -> -inf (ties survive)."""
    v, _ = torch.topk(logits, k)
    out = logits.clone()
    out[out < v[..., [-1]]] = -float("inf")
    return out


def sample_probs(logits_row, temperature=1.0, top_k=None):
    """cond_transformer_multi_view.py:200-211 -> probability vector the reference hands to multinomial."""
    l = logits_row / temperature
    if top_k is not None:
        l = top_k_logits(l, top_k)
    return F.softmax(l, dim=-1)


def sample_reference_loop(sd, geo, bev_indices, batch, steps, temperature=1.0, top_k=None, chooser=None):
    """The reference's O(L^2)-per-step loop (cond_transformer_multi_view.py:154-227) for `steps` tokens.

    chooser(probs, t) -> (B,) token ids (default greedy = `sample=False` branch, :216).
    Returns x (B,cam,tokens) with un-decoded positions = vocab (PAD) and the per-step logits rows.
    """
    B = bev_indices.shape[0]
    ncam, ntok = geo["num_cams"], geo["num_cam_tokens"]
    x = torch.full((B, ncam, ntok), geo["vocab_size"], dtype=torch.int64)
    rows = []
    fwd = geo["forward_shuffle_idx"]
    for t in range(steps):
        j = int(fwd[t])
        i, k = j // ntok, j % ntok
        logits = forward(sd, geo, x, bev_indices, batch, sampling=True).view(B, ncam, ntok, -1)[:, i, k]
        rows.append(logits)
        p = sample_probs(logits, temperature, top_k)
        ix = p.argmax(-1) if chooser is None else chooser(p, t)
        x[:, i, k] = ix
    return x, torch.stack(rows, 1)


class KVCacheDecoder:
    """KV-cache restatement of the sampling loop: prefill the cond tokens (they only see each
    other), then one row per generated token attending to keys [0, n_cond + t]; bias row added pre-scale."""

    def __init__(self, sd, geo):
        self.sd, self.geo = sd, geo
        self.bias = camera_bias(sd, geo)[0] if "camera_bias_emb" in sd else None

    def _layer_rows(self, i, x, K, V, row0):
        """x: (B,n,d) rows at sequence positions row0..row0+n-1 (already the block input); appends to K,V."""
        sd, geo = self.sd, self.geo
        p, nh = f"blocks.{i}", geo["num_heads"]
        d = x.shape[-1]
        x = F.layer_norm(x, (d,), sd[f"{p}.ln1.weight"], sd[f"{p}.ln1.bias"])
        B, n, _ = x.shape

        def proj(nm):
            return F.linear(x, sd[f"{p}.attention.{nm}.weight"], sd[f"{p}.attention.{nm}.bias"]).view(B, n, nh, d // nh).permute(0, 2, 1, 3)

        q, k, v = proj("query"), proj("key"), proj("value")
        K[i] = k if K[i] is None else torch.cat([K[i], k], 2)
        V[i] = v if V[i] is None else torch.cat([V[i], v], 2)
        nk = K[i].shape[2]
        s = q @ K[i].transpose(-1, -2)
        if self.bias is not None:
            s = s + self.bias[row0:row0 + n, :nk][None, None]
        s = s * (float(d // nh) ** -0.5)
        m = geo["attention_mask"][row0:row0 + n, :nk]
        s = s.masked_fill(m[None, None] == 0, float("-inf"))
        a = torch.softmax(s, -1) @ V[i]
        x = x + a.permute(0, 2, 1, 3).reshape(B, n, d)
        y = F.layer_norm(x, (d,), sd[f"{p}.ln2.weight"], sd[f"{p}.ln2.bias"])
        y = F.linear(F.gelu(F.linear(y, sd[f"{p}.mlp.0.weight"], sd[f"{p}.mlp.0.bias"])), sd[f"{p}.mlp.2.weight"], sd[f"{p}.mlp.2.bias"])
        return x + y

    def _logits(self, x_last):
        sd = self.sd
        d = x_last.shape[-1]
        return F.linear(F.layer_norm(x_last, (d,), sd["ln_f.weight"], sd["ln_f.bias"]), sd["head.weight"])

    def run(self, bev_indices, batch, steps, temperature=1.0, top_k=None, chooser=None, forced_tokens=None):
        sd, geo = self.sd, self.geo
        B = bev_indices.shape[0]
        ncam, ntok, nc = geo["num_cams"], geo["num_cam_tokens"], geo["num_cond_tokens"]
        pad_tok = torch.full((B, ncam, ntok), geo["vocab_size"], dtype=torch.int64)
        # embeddings of every position given a token grid; only rows <= current are ever used
        seq = embed(sd, geo, pad_tok, bev_indices, batch, sampling=True)
        K, V = [None] * geo["num_layers"], [None] * geo["num_layers"]
        x = seq[:, :nc]
        for i in range(geo["num_layers"]):
            x = self._layer_rows(i, x, K, V, 0)
        logits = self._logits(x[:, -1])
        out = pad_tok.clone()
        rows = []
        fwd = geo["forward_shuffle_idx"]
        for t in range(steps):
            rows.append(logits)
            p = sample_probs(logits, temperature, top_k)
            if forced_tokens is not None:
                ix = forced_tokens[:, t]
            else:
                ix = p.argmax(-1) if chooser is None else chooser(p, t)
            j = int(fwd[t])
            out[:, j // ntok, j % ntok] = ix
            if t == geo["num_img_tokens"] - 1:
                break
            seq = embed(sd, geo, out, bev_indices, batch, sampling=True)     # recompute (cheap); take the new row
            x = seq[:, nc + t: nc + t + 1]
            for i in range(geo["num_layers"]):
                x = self._layer_rows(i, x, K, V, nc + t)
            logits = self._logits(x[:, -1])
        return out, torch.stack(rows, 1)


def geo_from_config(cfg):
    """Collect the host artefacts the oracle needs from a (reference or bevgen_b200) GPTConfig object."""
    return dict(
        gpt_block_size=cfg.gpt_block_size, vocab_size=cfg.vocab_size, num_cams=cfg.num_cams,
        num_cam_tokens=cfg.num_cam_tokens, num_img_tokens=cfg.num_img_tokens, num_cond_tokens=cfg.num_cond_tokens,
        num_pad_tokens=cfg.num_pad_tokens, num_layers=cfg.num_layers, num_heads=cfg.num_heads,
        sparse_block_size=cfg.sparse_block_size, cam_latent_res=tuple(cfg.cam_latent_res),
        bev_latent_res=tuple(cfg.bev_latent_res), cam_res=tuple(cfg.cam_res),
        attention_mask=cfg.attention_mask, prob_matrix=cfg.prob_matrix,
        forward_shuffle_idx=cfg.forward_shuffle_idx, backward_shuffle_idx=cfg.backward_shuffle_idx)
