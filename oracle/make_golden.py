"""TEST INFRASTRUCTURE — mint tests/golden/*.npz by running the UNMODIFIED reference (imported from
/root/reference, see oracle/ref_import.py) on seeded synthetic weights/inputs (oracle/synth.py).

Run in the build container only:   python -m oracle.make_golden
The fixtures pin (a) the CPU restatements in oracle/ and (b) the product's host-side geometry.
"""
import sys
import zlib
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from oracle import ref_import, synth  # noqa: E402

OUT = ROOT / "tests" / "golden"

GPT_KW = dict(embd_pdrop=0., resid_pdrop=0., attn_pdrop=0., n_unmasked=0, num_cams=6, vocab_size=1024,
              cond_vocab_size=1024, hidden_size=1024, num_embed=1024, num_heads=16, num_layers=2,
              backend="deepspeed", sparse_block_size=16, window_len=32, cam_res=(256, 256),
              cam_latent_res=(16, 16), plot=False, causal_order=True, camera_bias=True, image_embed=True,
              bev_embed=True, bev_latent_res=(16, 16), density=1.0, cam_names="NUSCENES_CAMERAS", dataset="NUSCENES")

GPT_SMALL = {**GPT_KW, "hidden_size": 256, "num_embed": 256, "num_heads": 4, "vocab_size": 128, "cond_vocab_size": 128}
GPT_PADDED = {**GPT_SMALL, "cam_latent_res": (7, 9), "cam_res": (112, 144)}      # L=640 with 6 pad tokens

CONFIG_CASES = {
    "nusc6_16x16": GPT_KW,
    "nusc6_16x16_noncausal": {**GPT_KW, "causal_order": False},
    "nusc6_14x25": {**GPT_KW, "cam_latent_res": (14, 25), "cam_res": (224, 400)},
    "nusc6_7x9": GPT_PADDED,
    "nusc3_16x16": {**GPT_KW, "num_cams": 3, "cam_names": "NUSCENES_ABLATION_CAMERAS"},
    "argo3_16x16": {**GPT_KW, "num_cams": 3, "cam_names": "ARGOVERSE_FRONT_CAMERAS", "dataset": "ARGOVERSE"},
}


def crc(t):
    return np.int64(zlib.crc32(np.ascontiguousarray(t.detach().cpu().numpy()).tobytes()))


def golden_gptconfig():
    m = ref_import.stage2()
    for name, kw in CONFIG_CASES.items():
        torch.manual_seed(0)
        cfg = m.GPTConfig(**kw)
        layouts, allowed = cfg.get_mask()
        L = cfg.gpt_block_size
        rows = np.unique(np.concatenate([np.arange(0, L, 97), [0, 255, 256, 257, L - 1]])).astype(np.int64)
        rows = rows[rows < L]
        np.savez_compressed(
            OUT / f"gptconfig_{name}.npz",
            forward_shuffle_idx=cfg.forward_shuffle_idx.numpy().astype(np.int32),
            attention_mask_bits=np.packbits(cfg.attention_mask.numpy().astype(bool)),
            layout_bits=np.packbits(layouts.numpy().astype(bool)),
            layout_shape=np.array(layouts.shape),
            prob_rows=rows, prob_values=cfg.prob_matrix[rows].numpy(),
            prob_sum=np.float64(cfg.prob_matrix.sum().item()),
            prob_diag=torch.diagonal(cfg.prob_matrix).numpy(),
            sizes=np.array([cfg.gpt_block_size, cfg.num_cond_tokens, cfg.num_img_tokens, cfg.num_pad_tokens]))
        print("gptconfig", name, L)


def golden_outward_pattern():
    """mask_generator.outward_pattern (:131-214) itself: the intermediate tuple (allowed pattern, static block layout, block prior, padded
    prior) that multi_outward_pattern draws the per-head layouts from."""
    m = ref_import.stage2()
    from multi_view_generation.modules.transformer import mask_generator as mg
    for name in ("nusc6_16x16", "nusc6_16x16_noncausal", "nusc6_7x9"):
        torch.manual_seed(0)
        cfg = m.GPTConfig(**CONFIG_CASES[name])
        allowed, static_layout, prob_layout, prob_matrix = mg.outward_pattern(cfg)
        L = cfg.gpt_block_size
        rows = np.unique(np.concatenate([np.arange(0, L, 97), [0, 255, 256, 257, L - 1]])).astype(np.int64)
        rows = rows[rows < L]
        np.savez_compressed(
            OUT / f"outward_{name}.npz",
            allowed_bits=np.packbits(allowed[0].numpy().astype(bool)), allowed_shape=np.array(allowed.shape),
            allowed_dtype=str(allowed.dtype), static_bits=np.packbits(static_layout.numpy().astype(bool)),
            static_shape=np.array(static_layout.shape), static_dtype=str(static_layout.dtype),
            prob_layout=prob_layout.numpy(), prob_layout_dtype=str(prob_layout.dtype),
            prob_rows=rows, prob_values=prob_matrix[rows].numpy(), prob_sum=np.float64(prob_matrix.double().sum().item()),
            prob_dtype=str(prob_matrix.dtype), bias_sum=np.float64(mg.outward_pattern(cfg, return_camera_bias_matrix=True).double().sum().item()))
        print("outward", name, tuple(static_layout.shape), static_layout.dtype, prob_layout.dtype, prob_matrix.dtype, allowed.dtype)


def golden_vq():
    _, q = ref_import.stage1()
    for cb in ("normal", "default"):
        sd = synth.vqgan_state_dict(synth.vqgan_ddconfig(), seed=3, codebook=cb)
        vq = q.VectorQuantizer2(1024, 256, beta=0.25, legacy=True).eval()
        vq.embedding.weight.data.copy_(sd["quantize.embedding.weight"])
        z = synth.tensor_for("vq.z", (4, 256, 8, 8), seed=5, kind="embedding")
        if cb == "default":
            z = z * 1e-3
        with torch.no_grad():
            zq, loss, (_, _, idx) = vq(z)
            zq2 = vq.get_codebook_entry(idx, (4, 8, 8, 256))
        assert torch.equal(zq2, sd["quantize.embedding.weight"][idx].view(4, 8, 8, 256).permute(0, 3, 1, 2))
        np.savez_compressed(OUT / f"vq_{cb}.npz", idx=idx.numpy().astype(np.int32), zq_crc=crc(zq2), z_crc=crc(z),
                            zq_sample=zq2[0, :, 0, 0].numpy())
        print("vq", cb, idx[:8].tolist())


def _ref_vqgan(dd, sd):
    m, q = ref_import.stage1()
    enc, dec = m.Encoder(**dd).eval(), m.Decoder(**dd).eval()
    vq = q.VectorQuantizer2(1024, 256, beta=0.25, legacy=True).eval()
    qc = torch.nn.Conv2d(dd["z_channels"], 256, 1)
    pqc = torch.nn.Conv2d(256, dd["z_channels"], 1)
    enc.load_state_dict({k[len("encoder."):]: v for k, v in sd.items() if k.startswith("encoder.")})
    dec.load_state_dict({k[len("decoder."):]: v for k, v in sd.items() if k.startswith("decoder.")})
    vq.embedding.weight.data.copy_(sd["quantize.embedding.weight"])
    qc.load_state_dict({"weight": sd["quant_conv.weight"], "bias": sd["quant_conv.bias"]})
    pqc.load_state_dict({"weight": sd["post_quant_conv.weight"], "bias": sd["post_quant_conv.bias"]})
    return enc, dec, vq, qc, pqc


def golden_vqgan():
    cases = {
        # name: (ddconfig kwargs, batch, H, W)
        "small_rgb": (dict(in_channels=3, ch=64), 2, 64, 64),
        "small_bev": (dict(in_channels=7, ch=64), 1, 64, 64),
        "config1_rgb": (dict(in_channels=3, ch=128), 2, 128, 128),      # BASELINE.json configs[0]
    }
    for name, (kw, n, H, W) in cases.items():
        dd = synth.vqgan_ddconfig(**kw)
        sd = synth.vqgan_state_dict(dd, seed=1)
        enc, dec, vq, qc, pqc = _ref_vqgan(dd, sd)
        x = synth.image_batch(n, dd["in_channels"], H, W, seed=7)
        with torch.no_grad():                 # VQModel.encode / decode (vqgan.py:84-121), geometric_embedding=False
            h = qc(enc(x))
            quant, _, (_, _, idx) = vq(h)
            rec = dec(pqc(vq.get_codebook_entry(idx, (n, H // 16, W // 16, 256))))
        d = torch.cdist(h.permute(0, 2, 3, 1).reshape(-1, 256), sd["quantize.embedding.weight"])
        top2 = d.topk(2, largest=False).values
        np.savez_compressed(OUT / f"vqgan_{name}.npz", h=h.numpy(), idx=idx.numpy().astype(np.int32), rec=rec.numpy(),
                            x_crc=crc(x), min_gap=np.float64((top2[:, 1] - top2[:, 0]).min().item()))
        print("vqgan", name, tuple(rec.shape), "min top-2 gap", (top2[:, 1] - top2[:, 0]).min().item())


def golden_vqgan_geometric():
    """VQModel(geometric_embedding=True).encode / decode (vqgan.py:62-69,84-121; the default of configs/model/stage_1_cam.yaml) run through
    the reference's own LightningModule class on 2 scenes x 3 cameras of 64x64 pixels."""
    rv = ref_import.vqgan()
    dd = synth.vqgan_ddconfig(in_channels=3, ch=64, resolution=64)
    sd = synth.vqgan_state_dict(dd, seed=4, geometric=True)
    model = rv.VQModel(dd, None, 1024, 256, (64, 64), (4, 4), 256, geometric_embedding=True).eval()
    missing, unexpected = model.load_state_dict(sd, strict=True)
    g = torch.Generator().manual_seed(21)
    x = synth.image_batch(6, 3, 64, 64, seed=9)
    batch = {"intrinsics_inv": torch.randn(2, 3, 3, 3, generator=g) * 0.01, "extrinsics_inv": torch.randn(2, 3, 4, 4, generator=g)}
    pre = []
    hook = model.quant_conv.register_forward_hook(lambda mod, i, o: pre.append(o.detach()))
    with torch.no_grad():
        quant, _, (_, _, idx) = model.encode(x.clone(), batch)
        rec = model.decode(quant)
        plain, _, (_, _, idx_plain) = rv.VQModel.encode(_NoGeo(model), x.clone(), batch)
    hook.remove()
    assert not torch.equal(idx, idx_plain)              # the embedding really changes the tokens
    np.savez_compressed(OUT / "vqgan_geometric.npz", h=pre[0].numpy(), idx=idx.numpy().astype(np.int32), rec=rec.numpy(), x_crc=crc(x),
                        intrinsics_inv=batch["intrinsics_inv"].numpy(), extrinsics_inv=batch["extrinsics_inv"].numpy())
    print("vqgan geometric", tuple(rec.shape), "tokens changed by the embedding:", int((idx != idx_plain).sum()), "of", idx.numel())


class _NoGeo:
    """Attribute proxy that turns geometric_embedding off for one call of the reference's encode (to show the branch matters)."""
    def __init__(self, m):
        self._m = m

    def __getattr__(self, k):
        return False if k == "geometric_embedding" else getattr(self._m, k)


def _sizes(cfg):
    return dict(num_embed=cfg.num_embed, gpt_block_size=cfg.gpt_block_size, num_img_tokens=cfg.num_img_tokens,
                num_cond_tokens=cfg.num_cond_tokens, num_cams=cfg.num_cams, vocab_size=cfg.vocab_size,
                cond_vocab_size=cfg.cond_vocab_size, num_layers=cfg.num_layers)


def golden_gpt():
    m = ref_import.stage2()
    cases = {"small": (GPT_SMALL, 2), "padded": (GPT_PADDED, 2), "wide2": (GPT_KW, 1)}
    for name, (kw, B) in cases.items():
        torch.manual_seed(0)
        cfg = m.GPTConfig(**kw)
        model = m.GPT(cfg).eval()
        sd = synth.gpt_state_dict(_sizes(cfg), seed=2)
        missing, unexpected = model.load_state_dict(sd, strict=False)
        assert not unexpected and all("master_layout" in k or k == "bev_grid" for k in missing), (missing, unexpected)
        cam, bev, batch = synth.stage2_inputs(B, cfg.num_cams, cfg.num_cam_tokens, cfg.num_cond_tokens,
                                              cfg.vocab_size, cfg.cond_vocab_size, seed=4)
        hid = []
        hooks = [blk.register_forward_hook(lambda mod, i, o: hid.append(o[0].detach())) for blk in model.blocks]
        with torch.no_grad():
            logits_tf = model(cam.clone(), bev, batch, sampling=False)       # teacher-forced (pads last token)
            hid_tf = [h.clone() for h in hid]
            hid.clear()
            logits_s = model(cam.clone(), bev, batch, sampling=True)
        for h in hooks:
            h.remove()
        n = logits_tf.shape[1]
        rows = np.unique(np.concatenate([np.arange(0, n, 61), [0, 1, n - 2, n - 1]])).astype(np.int64)
        np.savez_compressed(OUT / f"gpt_{name}.npz", rows=rows, logits_tf=logits_tf[:, rows].numpy(),
                            logits_s=logits_s[:, rows].numpy(), logits_tf_mean=np.float64(logits_tf.double().mean().item()),
                            logits_tf_absmax=np.float64(logits_tf.abs().max().item()),
                            hidden0_rows=hid_tf[0][:, ::97].numpy(), hidden_last_rows=hid_tf[-1][:, ::97].numpy(),
                            cam_crc=crc(cam), B=np.int64(B))
        print("gpt", name, tuple(logits_tf.shape), float(logits_tf.abs().max()))
        if name == "small":
            # reference sampling loop (cond_transformer_multi_view.py:172-219), greedy, 4 steps, PAD-initialised x
            x = torch.full((B, cfg.num_cams, cfg.num_cam_tokens), cfg.vocab_size, dtype=torch.int64)
            rows_l, toks = [], []
            with torch.no_grad():
                for t in range(4):
                    j = int(cfg.forward_shuffle_idx[t])
                    i, k = j // cfg.num_cam_tokens, j % cfg.num_cam_tokens
                    lg = model(x, bev, batch, sampling=True).view(B, cfg.num_cams, cfg.num_cam_tokens, -1)[:, i, k]
                    probs = torch.softmax(lg, -1)
                    ix = probs.topk(1, dim=-1)[1].squeeze(-1)
                    x[:, i, k] = ix
                    rows_l.append(lg)
                    toks.append(ix)
            np.savez_compressed(OUT / "gpt_small_sample4.npz", logits=torch.stack(rows_l, 1).numpy(),
                                tokens=torch.stack(toks, 1).numpy().astype(np.int32))
            print("gpt sample4", torch.stack(toks, 1).tolist())


def golden_vqgan_config2():
    """BASELINE configs[1] shapes: ch=128 VQGAN at 256x256.  RGB: the first 8 images of the benchmark's 96-image batch (synth.image_batch(96,
    seed=100)); BEV tokenizer (7 channels): 2 scenes.  h / rec are stored sub-sampled (channel / pixel strides), idx in full."""
    for name, cin, n, seed_x in (("config2_rgb", 3, 8, 100), ("config2_bev", 7, 2, 300)):
        dd = synth.vqgan_ddconfig(in_channels=cin, ch=128)
        sd = synth.vqgan_state_dict(dd, seed=1)
        enc, dec, vq, qc, pqc = _ref_vqgan(dd, sd)
        x = synth.image_batch(96 if cin == 3 else 16, cin, 256, 256, seed=seed_x)[:n].contiguous()
        if cin == 7:
            x = (x > 0).float()                      # BEV segmentation maps are {0, 1} (SURVEY 8b batch schema)
        with torch.no_grad():
            h = qc(enc(x))
            quant, _, (_, _, idx) = vq(h)
            rec = dec(pqc(vq.get_codebook_entry(idx, (n, 16, 16, 256))))
        d = torch.cdist(h.permute(0, 2, 3, 1).reshape(-1, 256), sd["quantize.embedding.weight"])
        top2 = d.topk(2, largest=False).values
        np.savez_compressed(OUT / f"vqgan_{name}.npz", h_sub=h[:, ::8].numpy(), idx=idx.numpy().astype(np.int32),
                            rec_sub=rec[:, :, ::8, ::8].numpy(), rec_row=rec[:, :, 100].numpy(), x_crc=crc(x),
                            rec_absmax=np.float64(rec.abs().max().item()), rec_mean=np.float64(rec.double().mean().item()),
                            gap=(top2[:, 1] - top2[:, 0]).numpy(), min_gap=np.float64((top2[:, 1] - top2[:, 0]).min().item()))
        print("vqgan", name, tuple(rec.shape), "min top-2 gap", (top2[:, 1] - top2[:, 0]).min().item())


def _gpt_case_golden(m, name, kw, B, Bgen, layouts_from_reference=False, seed_in=4):
    """One reference GPT forward (teacher-forced + sampling) on the first B samples of a Bgen-sample seeded input batch."""
    torch.manual_seed(0)
    cfg = m.GPTConfig(**kw)
    model = m.GPT(cfg).eval()
    sd = synth.gpt_state_dict(_sizes(cfg), seed=2)
    missing, unexpected = model.load_state_dict(sd, strict=False)
    assert not unexpected and all("master_layout" in k or k == "bev_grid" for k in missing), (missing, unexpected)
    cam, bev, batch = synth.stage2_inputs(Bgen, cfg.num_cams, cfg.num_cam_tokens, cfg.num_cond_tokens, cfg.vocab_size, cfg.cond_vocab_size, seed=seed_in)
    cam, bev, batch = cam[:B].contiguous(), bev[:B].contiguous(), {k: v[:B].contiguous() for k, v in batch.items()}
    hid = []
    hooks = [blk.register_forward_hook(lambda mod, i, o: hid.append(o[0].detach())) for blk in model.blocks]
    with torch.no_grad():
        logits_tf = model(cam.clone(), bev, batch, sampling=False)
        hid_tf = [h.clone() for h in hid]
        hid.clear()
        logits_s = model(cam.clone(), bev, batch, sampling=True)
    for h in hooks:
        h.remove()
    n = logits_tf.shape[1]
    rows = np.unique(np.concatenate([np.arange(0, n, 61), [0, 1, n - 2, n - 1]])).astype(np.int64)
    extra = {}
    if layouts_from_reference:
        lay = torch.stack([blk.attention.sparse_self_attention.master_layout for blk in model.blocks])       # (layers, heads, nb, nb)
        extra = dict(layout_bits=np.packbits(lay.numpy().astype(bool)), layout_shape=np.array(lay.shape),
                     layout_density=np.float64(lay.float().mean().item()))
    np.savez_compressed(OUT / f"gpt_{name}.npz", rows=rows, logits_tf=logits_tf[:, rows].numpy(), logits_s=logits_s[:, rows].numpy(),
                        logits_tf_mean=np.float64(logits_tf.double().mean().item()), logits_tf_absmax=np.float64(logits_tf.abs().max().item()),
                        hidden0_rows=hid_tf[0][:, ::97].numpy(), hidden_last_rows=hid_tf[-1][:, ::97].numpy(), cam_crc=crc(cam), B=np.int64(B),
                        Bgen=np.int64(Bgen), **extra)
    print("gpt", name, tuple(logits_tf.shape), float(logits_tf.abs().max()), {k: float(v) for k, v in extra.items() if k == "layout_density"})


def golden_gpt_full():
    """BASELINE configs[2] model: 24 layers, d=1024, 16 heads, L=1792 - the first 2 samples of the benchmark's B=16 seeded batch."""
    _gpt_case_golden(ref_import.stage2(), "full24", {**GPT_KW, "num_layers": 24}, 2, 16, seed_in=0)


def golden_gpt_variants():
    """SURVEY 8f-2 / 8f-4 pinned to the reference: block-sparse layouts of density 0.25 and 0.5 DRAWN BY THE REFERENCE (per layer, per head;
    mask_generator.py:217-228) and stored with the logits; 3-camera Argoverse rig; AR decoder on the nuScenes-native 14x25 latents (L=2368,
    12 pad tokens)."""
    m = ref_import.stage2()
    _gpt_case_golden(m, "small_density25", {**GPT_SMALL, "density": 0.25}, 2, 2, layouts_from_reference=True)
    _gpt_case_golden(m, "small_density50", {**GPT_SMALL, "density": 0.5}, 1, 1, layouts_from_reference=True)
    _gpt_case_golden(m, "small_argo3", {**GPT_SMALL, "num_cams": 3, "cam_names": "ARGOVERSE_FRONT_CAMERAS", "dataset": "ARGOVERSE"}, 2, 2)
    _gpt_case_golden(m, "small_nusc14x25", {**GPT_SMALL, "cam_latent_res": (14, 25), "cam_res": (224, 400)}, 1, 1)


MASKGIT_DEPTH = 2


def golden_maskgit():
    """MaskGitTransformerMultiView.forward (logits + embed, masked and unmasked ids) and MaskGit.generate with a SelfCritic
    (muse_maskgit_pytorch.py:283-366,511-627), reference on CPU, seeded global RNG for the gumbel / critic noise."""
    mg = ref_import.muse()
    m = ref_import.stage2()
    kw, B = GPT_SMALL, 1
    cfg = m.GPTConfig(**kw)
    heads = cfg.num_heads
    tr = mg.MaskGitTransformerMultiView(num_tokens=cfg.vocab_size, dim=cfg.num_embed, seq_len=tuple(cfg.cam_latent_res), depth=MASKGIT_DEPTH,
                                        dim_head=64, heads=heads, ff_mult=4, cfg=cfg)
    model = mg.MaskGit(image_size=tuple(cfg.cam_latent_res), transformer=tr, self_token_critic=True).eval()
    sd = synth.maskgit_state_dict(_sizes(cfg), MASKGIT_DEPTH, heads, seed=3)
    crit = {k: sd.pop(k) for k in ("to_pred.weight", "to_pred.bias")}
    missing, unexpected = tr.load_state_dict(sd, strict=False)
    assert not unexpected and all(k.endswith(".beta") or k == "bev_grid" for k in missing), (missing, unexpected)
    model.token_critic.to_pred.load_state_dict({"weight": crit["to_pred.weight"], "bias": crit["to_pred.bias"]})
    cam, bev, batch = synth.stage2_inputs(B, cfg.num_cams, cfg.num_cam_tokens, cfg.num_cond_tokens, cfg.vocab_size, cfg.cond_vocab_size, seed=6)
    ids = cam.reshape(B * cfg.num_cams, cfg.num_cam_tokens).clone()
    g = torch.Generator().manual_seed(5)
    ids[torch.rand(ids.shape, generator=g) < 0.6] = tr.mask_id
    with torch.no_grad():
        logits, emb = tr(ids, return_embed=True, conditioning_token_ids=bev, batch=batch)
        guided = tr.forward_with_cond_scale(ids, conditioning_token_ids=bev, batch=batch, cond_scale=3.0)
        assert torch.equal(guided, logits)          # eval mode: the "null" pass equals the conditional pass (:341)
        torch.manual_seed(1234)
        gen = model.generate(cond_images=bev, fmap_size=tuple(cfg.cam_latent_res), batch=batch, timesteps=6)
    cols = np.arange(0, cfg.num_cam_tokens, 7)
    np.savez_compressed(OUT / "maskgit_small.npz", ids=ids.numpy().astype(np.int32), cols=cols, logits=logits[:, cols].numpy(),
                        embed=emb[:, cols].numpy(), logits_absmax=np.float64(logits.abs().max().item()),
                        logits_mean=np.float64(logits.double().mean().item()), generated=gen.numpy().astype(np.int32),
                        gen_seed=np.int64(1234), gen_steps=np.int64(6))
    print("maskgit", tuple(logits.shape), float(logits.abs().max()), "generated", tuple(gen.shape), int(gen.max()))
    import json
    json.dump({k: list(v.shape) for k, v in model.state_dict().items()}, open(OUT / "maskgit_small_state_dict_keys.json", "w"), indent=0,
              sort_keys=True)       # checkpoint-compatibility fixture: keys + shapes of MaskGit(...).state_dict()


def golden_topk():
    """Net2NetTransformer.top_k_logits (cond_transformer_multi_view.py:138-142) + softmax on fixed logits, incl. ties."""
    g = torch.Generator().manual_seed(11)
    logits = torch.randn(4, 1024, generator=g)
    logits[1, :200] = logits[1, 0]                      # 200-way tie above the k-th value boundary
    logits[2] = torch.round(logits[2] * 4) / 4          # heavy ties everywhere
    k = 100
    v, _ = torch.topk(logits, k)
    out = logits.clone()
    out[out < v[..., [-1]]] = -float("inf")
    probs = torch.softmax(out / 1.0, -1)
    np.savez_compressed(OUT / "topk.npz", logits=logits.numpy(), probs=probs.numpy(), k=np.int64(k),
                        kept=(out > -float("inf")).sum(-1).numpy())
    print("topk kept", (out > -float("inf")).sum(-1).tolist())


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "maskgit":
        golden_maskgit()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "vqgan_geometric":
        golden_vqgan_geometric()
        sys.exit(0)
    assert ref_import.available(), "run in the build container (needs /root/reference)"
    OUT.mkdir(parents=True, exist_ok=True)
    which = sys.argv[1:] or ["vq", "vqgan", "topk", "gptconfig", "gpt"]
    if "vq" in which:
        golden_vq()
    if "vqgan" in which:
        golden_vqgan()
    if "topk" in which:
        golden_topk()
    if "gptconfig" in which:
        golden_gptconfig()
    if "gpt" in which:
        golden_gpt()
    if "vqgan_config2" in which:
        golden_vqgan_config2()
    if "gpt_full" in which:
        golden_gpt_full()
    if "outward" in which:
        golden_outward_pattern()
    if "gpt_variants" in which:
        golden_gpt_variants()
