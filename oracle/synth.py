"""TEST INFRASTRUCTURE — deterministic synthetic weights and inputs keyed by state-dict key name.

There are no checkpoints offline, so parity is pinned on seeded random weights.
Every tensor is generated from ``crc32(key) ^ seed`` with its own ``torch.Generator`` so that the
golden-vector generator (which loads them into the *reference* modules) and the tests (which load
them into the oracle restatement and into the CUDA engine) see identical values regardless of
module construction order.
"""
import zlib
from collections import OrderedDict

import torch


def tensor_for(key: str, shape, seed: int = 0, kind: str = "auto") -> torch.Tensor:
    g = torch.Generator().manual_seed((zlib.crc32(key.encode()) ^ (seed * 0x9E3779B1)) & 0x7FFFFFFF)
    shape = tuple(int(s) for s in shape)
    r = torch.randn(shape, generator=g, dtype=torch.float32)
    if kind == "auto":
        if key.endswith("bias"):
            kind = "bias"
        elif ".norm" in key or "norm_out" in key or ".ln" in key or key.startswith("ln"):
            kind = "norm_w"
        else:
            kind = "weight"
    if kind == "bias":
        return 0.05 * r
    if kind == "norm_w":
        return 1.0 + 0.1 * r
    if kind == "embedding":
        return r
    if kind == "small":
        return 0.02 * r
    if kind == "bias_emb":
        return 0.1 * r
    # fan-in scaled weight (conv OIHW or linear OI)
    fan_in = 1
    for s in shape[1:]:
        fan_in *= s
    return r * (1.0 / max(fan_in, 1)) ** 0.5


def vqgan_ddconfig(in_channels=3, ch=128, ch_mult=(1, 1, 2, 2, 4), resolution=256, z_channels=256):
    return dict(double_z=False, z_channels=z_channels, resolution=resolution, in_channels=in_channels,
                out_ch=in_channels, ch=ch, ch_mult=list(ch_mult), num_res_blocks=2, attn_resolutions=[16],
                dropout=0.0)


def vqgan_keys(dd, n_embed=1024, embed_dim=256):
    """(key, shape) list of a VQModel state dict (reference key names: SURVEY.md §8b)."""
    out = []

    def conv(name, cin, cout, k):
        out.append((f"{name}.weight", (cout, cin, k, k)))
        out.append((f"{name}.bias", (cout,)))

    def norm(name, c):
        out.append((f"{name}.weight", (c,)))
        out.append((f"{name}.bias", (c,)))

    def res(name, cin, cout):
        norm(f"{name}.norm1", cin)
        conv(f"{name}.conv1", cin, cout, 3)
        norm(f"{name}.norm2", cout)
        conv(f"{name}.conv2", cout, cout, 3)
        if cin != cout:
            conv(f"{name}.nin_shortcut", cin, cout, 1)

    def attn(name, c):
        norm(f"{name}.norm", c)
        for p in ("q", "k", "v", "proj_out"):
            conv(f"{name}.{p}", c, c, 1)

    ch, mult, nres = dd["ch"], list(dd["ch_mult"]), dd["num_res_blocks"]
    nlev = len(mult)
    # encoder (model.py:342-404)
    conv("encoder.conv_in", dd["in_channels"], ch, 3)
    cur = dd["resolution"]
    in_mult = [1] + mult
    bi = ch
    for l in range(nlev):
        bi, bo = ch * in_mult[l], ch * mult[l]
        for b in range(nres):
            res(f"encoder.down.{l}.block.{b}", bi, bo)
            bi = bo
            if cur in dd["attn_resolutions"]:
                attn(f"encoder.down.{l}.attn.{b}", bi)
        if l != nlev - 1:
            conv(f"encoder.down.{l}.downsample.conv", bi, bi, 3)
            cur //= 2
    res("encoder.mid.block_1", bi, bi)
    attn("encoder.mid.attn_1", bi)
    res("encoder.mid.block_2", bi, bi)
    norm("encoder.norm_out", bi)
    conv("encoder.conv_out", bi, dd["z_channels"], 3)
    # decoder (model.py:436-504)
    bi = ch * mult[-1]
    cur = dd["resolution"] // 2 ** (nlev - 1)
    conv("decoder.conv_in", dd["z_channels"], bi, 3)
    res("decoder.mid.block_1", bi, bi)
    attn("decoder.mid.attn_1", bi)
    res("decoder.mid.block_2", bi, bi)
    for l in reversed(range(nlev)):
        bo = ch * mult[l]
        for b in range(nres + 1):
            res(f"decoder.up.{l}.block.{b}", bi, bo)
            bi = bo
            if cur in dd["attn_resolutions"]:
                attn(f"decoder.up.{l}.attn.{b}", bi)
        if l != 0:
            conv(f"decoder.up.{l}.upsample.conv", bi, bi, 3)
            cur *= 2
    norm("decoder.norm_out", bi)
    conv("decoder.conv_out", bi, dd["out_ch"], 3)
    out.append(("quantize.embedding.weight", (n_embed, embed_dim)))
    conv("quant_conv", dd["z_channels"], embed_dim, 1)
    conv("post_quant_conv", embed_dim, dd["z_channels"], 1)
    return out


def vqgan_state_dict(dd, seed=0, n_embed=1024, embed_dim=256, codebook="normal", geometric=False):
    sd = OrderedDict()
    if geometric:      # VQModel(geometric_embedding=True): 1x1 convs of the ray / camera-centre embedding (vqgan.py:68-69), cam_emd_dim = z_channels
        for k in ("img_embed.weight", "cam_embed.weight"):
            sd[k] = tensor_for(k, (dd["z_channels"], 4, 1, 1), seed, "weight")
    for k, shp in vqgan_keys(dd, n_embed, embed_dim):
        if k == "quantize.embedding.weight":
            if codebook == "normal":      # separated codes: bit-exact argmin is well defined
                sd[k] = tensor_for(k, shp, seed, "embedding")
            else:                          # reference default init U(-1/n, 1/n) (quantize.py:232)
                g = torch.Generator().manual_seed(seed + 17)
                sd[k] = (torch.rand(shp, generator=g) * 2 - 1) / n_embed
        else:
            sd[k] = tensor_for(k, shp, seed)
    return sd


def gpt_keys(c):
    """(key, shape, kind) for the reference GPT state dict (mingpt_sparse.py:270-311). c: dict of sizes."""
    d, L = c["num_embed"], c["gpt_block_size"]
    out = [("x_pos_emb", (1, c["num_img_tokens"], d), "small"), ("cond_pos_emb", (1, c["num_cond_tokens"], d), "small")]
    if c.get("bev_embed", True):
        out.append(("bev_cam_pos_emb", (1, c["num_cams"], c["num_cond_tokens"], d), "small"))
    if c.get("camera_bias", True):
        out.append(("camera_bias_emb", (1, L * (L + 1) // 2), "bias_emb"))
    out += [("x_tok_emb.weight", (c["vocab_size"] + 1, d), "small"), ("cond_tok_emb.weight", (c["cond_vocab_size"], d), "small")]
    for i in range(c["num_layers"]):
        p = f"blocks.{i}"
        out += [(f"{p}.ln1.weight", (d,), "norm_w"), (f"{p}.ln1.bias", (d,), "bias"),
                (f"{p}.ln2.weight", (d,), "norm_w"), (f"{p}.ln2.bias", (d,), "bias")]
        for n in ("query", "key", "value"):
            out += [(f"{p}.attention.{n}.weight", (d, d), "small"), (f"{p}.attention.{n}.bias", (d,), "bias")]
        out += [(f"{p}.mlp.0.weight", (4 * d, d), "small"), (f"{p}.mlp.0.bias", (4 * d,), "bias"),
                (f"{p}.mlp.2.weight", (d, 4 * d), "small"), (f"{p}.mlp.2.bias", (d,), "bias")]
    out += [("ln_f.weight", (d,), "norm_w"), ("ln_f.bias", (d,), "bias"), ("head.weight", (c["vocab_size"], d), "small")]
    if c.get("image_embed", True):
        out += [("img_embed.weight", (d, 4, 1, 1), "weight"), ("cam_embed.weight", (d, 4, 1, 1), "weight")]
    if c.get("bev_embed", True):
        out += [("bev_embed.weight", (d, 2, 1, 1), "weight"), ("bev_embed.bias", (d,), "bias")]
    return out


def gpt_state_dict(c, seed=0):
    sd = OrderedDict()
    for k, shp, kind in gpt_keys(c):
        sd[k] = tensor_for(k, shp, seed, kind)
    return sd


def stage2_inputs(batch, num_cams=6, cam_tokens=256, cond_tokens=256, vocab=1024, cond_vocab=1024, seed=0):
    """Seeded random tokens + camera matrices as in the reference's own smoke script (test_masks.py:20-21)."""
    g = torch.Generator().manual_seed(1000 + seed)
    cam = torch.randint(0, vocab, (batch, num_cams, cam_tokens), generator=g)
    bev = torch.randint(0, cond_vocab, (batch, cond_tokens), generator=g)
    I_inv = torch.randn(batch, num_cams, 3, 3, generator=g)
    E_inv = torch.randn(batch, num_cams, 4, 4, generator=g)
    return cam, bev, {"intrinsics_inv": I_inv, "extrinsics_inv": E_inv}


def image_batch(n, c=3, h=256, w=256, seed=0):
    g = torch.Generator().manual_seed(2000 + seed)
    return torch.randn(n, c, h, w, generator=g)


def maskgit_keys(c, depth, heads, dim_head=64, ff_mult=4):
    """(key, shape, kind) of the reference MaskGitTransformerMultiView state dict (muse_maskgit_pytorch.py:90-129,171-247), the unused
    `norm.*` / `self_cond_to_init_embed.*` members included so that load_state_dict(strict) finds every parameter."""
    d, L, inner, ffi = c["num_embed"], c["gpt_block_size"], heads * dim_head, int(c["num_embed"] * ff_mult * 2 / 3)
    out = [("token_emb.weight", (c["vocab_size"] + 1, d), "small"), ("pos_emb.weight", (c["num_img_tokens"], d), "small"),
           ("cond_token_emb.weight", (c["cond_vocab_size"], d), "small"), ("cond_pos_emb.weight", (c["num_cond_tokens"], d), "small")]

    def ff(p):
        return [(f"{p}.0.gamma", (d,), "norm_w"), (f"{p}.1.weight", (2 * ffi, d), "weight"), (f"{p}.3.gamma", (ffi,), "norm_w"),
                (f"{p}.4.weight", (d, ffi), "weight")]
    for i in range(depth):
        for a in (0, 1):
            p = f"transformer_blocks.layers.{i}.{a}"
            out += [(f"{p}.norm.gamma", (d,), "norm_w"), (f"{p}.null_kv", (2, heads, 1, dim_head), "embedding"),
                    (f"{p}.to_q.weight", (inner, d), "weight"), (f"{p}.to_kv.weight", (2 * inner, d), "weight"),
                    (f"{p}.q_scale", (dim_head,), "norm_w"), (f"{p}.k_scale", (dim_head,), "norm_w"), (f"{p}.to_out.weight", (d, inner), "weight")]
        out += ff(f"transformer_blocks.layers.{i}.2")
    out += [("transformer_blocks.norm.gamma", (d,), "norm_w"), ("norm.gamma", (d,), "norm_w"), ("to_logits.weight", (c["vocab_size"], d), "small")]
    out += ff("self_cond_to_init_embed")
    out += [("img_embed.weight", (d, 4, 1, 1), "weight"), ("cam_embed.weight", (d, 4, 1, 1), "weight"),
            ("bev_embed.weight", (d, 2, 1, 1), "weight"), ("bev_embed.bias", (d,), "bias"),
            ("bev_cam_pos_emb", (1, c["num_cams"], c["num_cond_tokens"], d), "small"), ("camera_bias_emb", (1, L * (L + 1) // 2), "bias_emb")]
    return out


def maskgit_state_dict(c, depth, heads, seed=0, critic=True):
    """Seeded MaskGit transformer weights (keys as in the reference) plus, with `critic`, the SelfCritic head `to_pred.*`
    (muse_maskgit_pytorch.py:371-375)."""
    sd = OrderedDict()
    for k, shp, kind in maskgit_keys(c, depth, heads):
        sd[k] = tensor_for(k, shp, seed, kind)
    if critic:
        sd["to_pred.weight"] = tensor_for("to_pred.weight", (1, c["num_embed"]), seed, "weight")
        sd["to_pred.bias"] = tensor_for("to_pred.bias", (1,), seed, "bias")
    return sd
