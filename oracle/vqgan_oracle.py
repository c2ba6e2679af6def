"""TEST INFRASTRUCTURE — CPU fp32 restatement of the reference's stage-1 VQGAN inference path.

Functional (state-dict driven) restatement of
  Encoder.forward            multi_view_generation/modules/stage1/model.py:406-433
  Decoder.forward            model.py:506-537
  ResnetBlock.forward        model.py:117-137      AttnBlock.forward  model.py:168-192
  Downsample / Upsample      model.py:68-75 / 49-53   Normalize = GroupNorm(32, eps=1e-6) model.py:34-35
  VQModel.encode / decode    modules/stage1/vqgan.py:84-121 (geometric_embedding=False branch)
  VectorQuantizer2.forward   modules/stage1/quantize.py:271-312 ; get_codebook_entry :314-329
Pinned against the unmodified reference modules by tests/golden/vqgan_*.pt (oracle/make_golden.py).
Not shipped, never on the product path.
"""
import torch
import torch.nn.functional as F


def _gn(x, sd, name):
    return F.group_norm(x, 32, sd[f"{name}.weight"], sd[f"{name}.bias"], eps=1e-6)


def _swish(x):
    return x * torch.sigmoid(x)


def _conv(x, sd, name, stride=1, padding=0):
    return F.conv2d(x, sd[f"{name}.weight"], sd[f"{name}.bias"], stride=stride, padding=padding)


def resnet_block(x, sd, name):
    h = _conv(_swish(_gn(x, sd, f"{name}.norm1")), sd, f"{name}.conv1", padding=1)
    h = _conv(_swish(_gn(h, sd, f"{name}.norm2")), sd, f"{name}.conv2", padding=1)
    if f"{name}.nin_shortcut.weight" in sd:
        x = _conv(x, sd, f"{name}.nin_shortcut")
    return x + h


def attn_block(x, sd, name):
    h = _gn(x, sd, f"{name}.norm")                       # no swish here (model.py:170)
    q, k, v = (_conv(h, sd, f"{name}.{p}") for p in ("q", "k", "v"))
    b, c, hh, ww = q.shape
    q = q.reshape(b, c, hh * ww).permute(0, 2, 1)
    k = k.reshape(b, c, hh * ww)
    w = torch.bmm(q, k) * (int(c) ** -0.5)
    w = torch.softmax(w, dim=2)
    v = v.reshape(b, c, hh * ww)
    h = torch.bmm(v, w.permute(0, 2, 1)).reshape(b, c, hh, ww)
    return x + _conv(h, sd, f"{name}.proj_out")


def downsample(x, sd, name):
    return _conv(F.pad(x, (0, 1, 0, 1)), sd, f"{name}.conv", stride=2)       # asymmetric pad, model.py:70-72


def upsample(x, sd, name):
    return _conv(F.interpolate(x, scale_factor=2.0, mode="nearest"), sd, f"{name}.conv", padding=1)


def _levels(sd, prefix):
    l = 0
    while f"{prefix}.{l}.block.0.norm1.weight" in sd:
        l += 1
    return l


def encoder(x, sd, num_res_blocks=2):
    h = _conv(x, sd, "encoder.conv_in", padding=1)
    nlev = _levels(sd, "encoder.down")
    for l in range(nlev):
        for b in range(num_res_blocks):
            h = resnet_block(h, sd, f"encoder.down.{l}.block.{b}")
            if f"encoder.down.{l}.attn.{b}.norm.weight" in sd:
                h = attn_block(h, sd, f"encoder.down.{l}.attn.{b}")
        if l != nlev - 1:
            h = downsample(h, sd, f"encoder.down.{l}.downsample")
    h = resnet_block(h, sd, "encoder.mid.block_1")
    h = attn_block(h, sd, "encoder.mid.attn_1")
    h = resnet_block(h, sd, "encoder.mid.block_2")
    return _conv(_swish(_gn(h, sd, "encoder.norm_out")), sd, "encoder.conv_out", padding=1)


def decoder(z, sd, num_res_blocks=2):
    h = _conv(z, sd, "decoder.conv_in", padding=1)
    h = resnet_block(h, sd, "decoder.mid.block_1")
    h = attn_block(h, sd, "decoder.mid.attn_1")
    h = resnet_block(h, sd, "decoder.mid.block_2")
    nlev = _levels(sd, "decoder.up")
    for l in reversed(range(nlev)):
        for b in range(num_res_blocks + 1):
            h = resnet_block(h, sd, f"decoder.up.{l}.block.{b}")
            if f"decoder.up.{l}.attn.{b}.norm.weight" in sd:
                h = attn_block(h, sd, f"decoder.up.{l}.attn.{b}")
        if l != 0:
            h = upsample(h, sd, f"decoder.up.{l}.upsample")
    return _conv(_swish(_gn(h, sd, "decoder.norm_out")), sd, "decoder.conv_out", padding=1)


def vq_nearest(z_nchw, codebook):
    """quantize.py:276-285: d = |z|^2 + |e|^2 - 2 z.e ; argmin ; gather.  Returns (z_q NCHW, idx (N*H*W,), d)."""
    z = z_nchw.permute(0, 2, 3, 1).contiguous()
    zf = z.view(-1, codebook.shape[1])
    d = torch.sum(zf ** 2, dim=1, keepdim=True) + torch.sum(codebook ** 2, dim=1) - 2 * torch.einsum(
        "bd,dn->bn", zf, codebook.t())
    idx = torch.argmin(d, dim=1)
    z_q = codebook[idx].view(z.shape).permute(0, 3, 1, 2).contiguous()
    return z_q, idx, d


def ray_embedding(sd, batch, cam_res, lat_res):
    """vqgan.py:62-69,87-109 (geometric_embedding=True): per camera and latent pixel the L2-normalised difference between the embedded
    viewing ray  img_embed(E_inv (I_inv (x*W, y*H, 1); 1))  and the embedded camera centre  cam_embed(E_inv[:, 3])  -> (b*cam, d, h, w).
    Note the stage-1 plane scales x by the image WIDTH and y by the HEIGHT (cam_res = (H, W)), unlike stage 2 (mingpt_sparse.py)."""
    fh, fw = lat_res
    xs, ys = torch.linspace(0, 1, fw), torch.linspace(0, 1, fh)
    gx, gy = torch.meshgrid((xs, ys), indexing="xy")
    pix = torch.stack([gx * cam_res[1], gy * cam_res[0], torch.ones_like(gx)], 0).reshape(3, fh * fw)      # 3 (h w)
    I_inv = batch["intrinsics_inv"].float().reshape(-1, 3, 3)
    E_inv = batch["extrinsics_inv"].float().reshape(-1, 4, 4)
    d = sd["img_embed.weight"].shape[0]
    cam = I_inv @ pix                                                   # n 3 (h w)
    cam = torch.cat([cam, torch.ones_like(cam[:, :1])], 1)              # n 4 (h w)
    ray = E_inv @ cam                                                   # n 4 (h w)
    d_embed = torch.einsum("dk,nkp->ndp", sd["img_embed.weight"].reshape(d, 4), ray)
    c_embed = torch.einsum("dk,nk->nd", sd["cam_embed.weight"].reshape(d, 4), E_inv[:, :, 3])
    e = d_embed - c_embed[:, :, None]
    e = e / (e.norm(dim=1, keepdim=True) + 1e-7)
    return e.reshape(-1, d, fh, fw)


def encode(x, sd, batch=None, cam_res=None):
    """VQModel.encode (vqgan.py:84-116) -> (quant NCHW, idx flat int64, pre-quant h).  With `img_embed.weight` in sd and a batch the
    geometric_embedding=True branch adds the ray embedding to the encoder output before quant_conv."""
    h = encoder(x, sd)
    if batch is not None and "img_embed.weight" in sd:
        h = h + ray_embedding(sd, batch, cam_res, h.shape[-2:])
    h = _conv(h, sd, "quant_conv")
    z_q, idx, _ = vq_nearest(h, sd["quantize.embedding.weight"])
    return z_q, idx, h


def get_codebook_entry(idx, shape_bhwc, sd):
    """quantize.py:314-329."""
    return sd["quantize.embedding.weight"][idx].view(shape_bhwc).permute(0, 3, 1, 2).contiguous()


def decode(quant, sd):
    """VQModel.decode (vqgan.py:118-121)."""
    return decoder(_conv(quant, sd, "post_quant_conv"), sd)


def denormalize(x):
    """bev_utils/util.py:97-118 (keep_tensor=True): x*std+mean per channel, clamp to [0,1]."""
    mean = torch.tensor([0.4265, 0.4489, 0.4769]).view(1, 3, 1, 1)
    std = torch.tensor([0.2053, 0.2206, 0.2578]).view(1, 3, 1, 1)
    return torch.clamp(x * std + mean, 0, 1)
