"""TEST INFRASTRUCTURE — CPU fp32 restatement of the reference's MaskGit stage-2 variant.

Functional (state-dict driven) restatement of, under /root/reference/multi_view_generation/modules/stage2/muse_maskgit_pytorch.py:
  TransformerMultiView.forward      :283-366   embeddings (token + ray + position; BEV context), camera bias, blocks, logits
  Attention.forward                 :116-169   LayerNorm(gamma) -> q / kv Linear -> null key/value -> l2norm(q), l2norm(k) * learned scales
                                               -> softmax(8 * q.k + bias slice) -> to_out          (self: bias[nc:, nc:], cross: bias[nc:, :nc])
  FeedForward / GEGLU               :72-88     LayerNorm -> Linear(d, 2f) -> gate * gelu(x) -> LayerNorm(f) -> Linear(f, d),  f = int(d*4*2/3)
  TransformerBlocks.forward         :192-202   x += self(x); x += cross(x, context); x += ff(x); final LayerNorm
  forward_with_cond_scale           :262-281   in eval mode `cond_drop_prob` is ignored (:341), so the "null" pass equals the conditional
                                               pass and the guided logits equal the plain logits; the oracle runs ONE pass
  SelfCritic                        :371-381   scores = Linear(d, 1)(embed)
  MaskGit.generate                  :511-627   cosine schedule, top-k (1 - thres) filter, gumbel sampling, critic scores + annealed noise
`noise(kind, step, shape)` supplies the uniform random numbers so that the CUDA engine and this oracle can replay the same draws.
Pinned against the reference by tests/golden/maskgit_small.npz. Not shipped, never on the product path.
"""
import math

import torch
import torch.nn.functional as F

from oracle import gpt_oracle


def _ln(x, gamma):
    return F.layer_norm(x, x.shape[-1:], gamma, torch.zeros_like(gamma))


def embed(sd, geo, ids, cond_ids, batch):
    """-> (x [b, n_img, d], context [b, nc, d]); ids (b*cam, hw) may hold the mask id = vocab_size."""
    I_inv, E_inv = batch["intrinsics_inv"].float(), batch["extrinsics_inv"].float()
    b, ncam = I_inv.shape[:2]
    h, w = geo["cam_latent_res"]
    d = sd["token_emb.weight"].shape[1]
    x = sd["token_emb.weight"][ids.reshape(b, ncam, h * w)]
    plane = gpt_oracle.generate_grid(h, w)[None]
    plane[:, :, 0] *= geo["cam_res"][0]
    plane[:, :, 1] *= geo["cam_res"][1]
    c_embed = torch.einsum("dk,bnk->bnd", sd["cam_embed.weight"].reshape(d, 4), E_inv[..., -1])
    cam = F.pad(I_inv @ plane.reshape(1, 1, 3, h * w), (0, 0, 0, 1), value=1)
    ray = E_inv @ cam
    e = torch.einsum("dk,bnkp->bnpd", sd["img_embed.weight"].reshape(d, 4), ray) - c_embed[:, :, None, :]
    x = x + e / (e.norm(dim=-1, keepdim=True) + 1e-7)
    x = x.reshape(b, ncam * h * w, d) + sd["pos_emb.weight"]
    ctx = sd["cond_token_emb.weight"][cond_ids]
    bh, bw = geo["bev_latent_res"]
    g = gpt_oracle.bev_grid(bh, bw)[:2].reshape(2, bh * bw)
    grid_embed = (sd["bev_embed.weight"].reshape(d, 2) @ g).t() + sd["bev_embed.bias"]
    ctx = ctx + (grid_embed[None] - (sd["bev_cam_pos_emb"] + c_embed[:, :, None, :]).sum(1))
    return x, ctx + sd["cond_pos_emb.weight"]


def attention(sd, p, x, kv_in, bias, heads):
    b, n, _ = x.shape
    xn = _ln(x, sd[f"{p}.norm.gamma"])
    src = xn if kv_in is None else kv_in
    q = (xn @ sd[f"{p}.to_q.weight"].t()) * 8.0
    k, v = (src @ sd[f"{p}.to_kv.weight"].t()).chunk(2, -1)
    split = lambda t: t.reshape(b, t.shape[1], heads, -1).transpose(1, 2)
    q, k, v = split(q), split(k), split(v)
    nk, nv = sd[f"{p}.null_kv"]
    k = torch.cat([nk[None].expand(b, -1, -1, -1), k], 2)
    v = torch.cat([nv[None].expand(b, -1, -1, -1), v], 2)
    q = F.normalize(q, dim=-1) * sd[f"{p}.q_scale"]
    k = F.normalize(k, dim=-1) * sd[f"{p}.k_scale"]
    sim = q @ k.transpose(-1, -2) * 8.0
    if bias is not None:
        sim = sim + F.pad(bias, (1, 0), value=0.0)
    o = sim.softmax(-1) @ v
    return o.transpose(1, 2).reshape(b, n, -1) @ sd[f"{p}.to_out.weight"].t()


def feed_forward(sd, p, x):
    h = _ln(x, sd[f"{p}.0.gamma"]) @ sd[f"{p}.1.weight"].t()
    a, gate = h.chunk(2, -1)
    return _ln(gate * F.gelu(a), sd[f"{p}.3.gamma"]) @ sd[f"{p}.4.weight"].t()


def forward(sd, geo, ids, cond_ids, batch, depth, heads):
    """-> (logits [(b cam), hw, vocab], embed [(b cam), hw, d]) of MaskGitTransformerMultiView.forward(return_embed=True) in eval mode."""
    x, ctx = embed(sd, geo, ids, cond_ids, batch)
    nc = geo["num_cond_tokens"]
    bias = gpt_oracle.camera_bias({"camera_bias_emb": sd["camera_bias_emb"]}, geo)[0] if "camera_bias_emb" in sd else None
    for i in range(depth):
        p = f"transformer_blocks.layers.{i}"
        x = attention(sd, f"{p}.0", x, None, None if bias is None else bias[nc:, nc:], heads) + x
        x = attention(sd, f"{p}.1", x, ctx, None if bias is None else bias[nc:, :nc], heads) + x
        x = feed_forward(sd, f"{p}.2", x) + x
    emb = _ln(x, sd["transformer_blocks.norm.gamma"])
    logits = emb @ sd["to_logits.weight"].t()
    b = x.shape[0]
    ncam, hw = geo["num_cams"], geo["num_cam_tokens"]
    return logits.reshape(b * ncam, hw, -1), emb.reshape(b * ncam, hw, -1)


def top_k(logits, thres=0.9):
    k = math.ceil((1 - thres) * logits.shape[-1])
    val, ind = logits.topk(k, dim=-1)
    return torch.full_like(logits, float("-inf")).scatter_(2, ind, val)


def generate(sd, geo, cond_ids, batch, depth, heads, noise, timesteps=18, temperature=1.0, topk_filter_thres=0.9, critic_noise_scale=1.0,
             init_ids=None, use_critic=True, trace=None):
    """MaskGit.generate (:511-627) with a SelfCritic.  noise(kind, step, shape) -> uniform(0,1) tensor, kind in {"gumbel", "critic"}."""
    mask_id = geo["vocab_size"]
    b = cond_ids.shape[0]
    ncam, hw = geo["num_cams"], geo["num_cam_tokens"]
    shape = (b * ncam, hw)
    scores = torch.zeros(shape, dtype=torch.float32)
    ids = torch.full(shape, mask_id, dtype=torch.long)
    init_mask = None if init_ids is None else init_ids != mask_id
    for step, (t, until_x0) in enumerate(zip(torch.linspace(0, 1, timesteps), reversed(range(timesteps)))):
        n_mask = max(int((torch.cos(t * math.pi * 0.5) * hw).item()), 1)
        ids = ids.scatter(1, scores.topk(n_mask, dim=-1).indices, mask_id)
        if init_ids is not None:
            ids[init_mask] = init_ids[init_mask]
        logits, _ = forward(sd, geo, ids, cond_ids, batch, depth, heads)
        if trace is not None:
            trace.append((ids.clone(), logits.clone()))
        temp = temperature * (until_x0 / timesteps)
        u = noise("gumbel", step, logits.shape)
        g = -torch.log((-torch.log(u.clamp(min=1e-20))).clamp(min=1e-20))
        pred = (top_k(logits, topk_filter_thres) / max(temp, 1e-10) + g).argmax(-1)
        is_mask = ids == mask_id
        ids = torch.where(is_mask, pred, ids)
        if use_critic:
            _, emb = forward(sd, geo, ids, cond_ids, batch, depth, heads)
            scores = (emb @ sd["to_pred.weight"].t() + sd["to_pred.bias"])[..., 0]
            scores = scores + (noise("critic", step, scores.shape) - 0.5) * critic_noise_scale * (until_x0 / timesteps)
        else:
            scores = 1 - logits.softmax(-1).gather(2, pred[..., None])[..., 0]
            scores = scores.masked_fill(~is_mask, -1e5)
    return ids.reshape(b * ncam, *geo["cam_latent_res"])
