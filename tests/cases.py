"""Shared config dictionaries for tests (kept identical to oracle/make_golden.py's cases)."""
GPT_KW = dict(embd_pdrop=0., resid_pdrop=0., attn_pdrop=0., n_unmasked=0, num_cams=6, vocab_size=1024,
              cond_vocab_size=1024, hidden_size=1024, num_embed=1024, num_heads=16, num_layers=2,
              backend="deepspeed", sparse_block_size=16, window_len=32, cam_res=(256, 256),
              cam_latent_res=(16, 16), plot=False, causal_order=True, camera_bias=True, image_embed=True,
              bev_embed=True, bev_latent_res=(16, 16), density=1.0, cam_names="NUSCENES_CAMERAS", dataset="NUSCENES")
GPT_SMALL = {**GPT_KW, "hidden_size": 256, "num_embed": 256, "num_heads": 4, "vocab_size": 128, "cond_vocab_size": 128}
GPT_PADDED = {**GPT_SMALL, "cam_latent_res": (7, 9), "cam_res": (112, 144)}
CONFIG_CASES = {
    "nusc6_16x16": GPT_KW,
    "nusc6_16x16_noncausal": {**GPT_KW, "causal_order": False},
    "nusc6_14x25": {**GPT_KW, "cam_latent_res": (14, 25), "cam_res": (224, 400)},
    "nusc6_7x9": GPT_PADDED,
    "nusc3_16x16": {**GPT_KW, "num_cams": 3, "cam_names": "NUSCENES_ABLATION_CAMERAS"},
    "argo3_16x16": {**GPT_KW, "num_cams": 3, "cam_names": "ARGOVERSE_FRONT_CAMERAS", "dataset": "ARGOVERSE"},
}
GPT_CASES = {"small": (GPT_SMALL, 2), "padded": (GPT_PADDED, 2), "wide2": (GPT_KW, 1)}
# goldens of oracle/make_golden.py golden_gpt_full / golden_gpt_variants: name -> (config, B used, B generated, input seed, reference-drawn layouts?)
GPT_FULL = {**GPT_KW, "num_layers": 24}
GPT_VARIANTS = {
    "small_density25": ({**GPT_SMALL, "density": 0.25}, 2, 2, 4, True),
    "small_density50": ({**GPT_SMALL, "density": 0.5}, 1, 1, 4, True),
    "small_argo3": ({**GPT_SMALL, "num_cams": 3, "cam_names": "ARGOVERSE_FRONT_CAMERAS", "dataset": "ARGOVERSE"}, 2, 2, 4, False),
    "small_nusc14x25": ({**GPT_SMALL, "cam_latent_res": (14, 25), "cam_res": (224, 400)}, 1, 1, 4, False),
    "full24": (GPT_FULL, 2, 16, 0, False),
}


def gpt_variant_inputs(name, synth, GPTConfig, B=None):
    """Config + seeded weights + the golden's inputs (first B samples of a Bgen-sample seeded batch); layouts come from the golden file."""
    kw, Bg, Bgen, seed_in, _ = GPT_VARIANTS[name]
    B = Bg if B is None else B
    if kw.get("density", 1.0) < 1.0:
        kw = {**kw}       # the product draws its own layouts at construction; tests replace them by the reference-drawn ones
    cfg = GPTConfig(**kw)
    sd = synth.gpt_state_dict(gpt_sizes(cfg), seed=2)
    cam, bev, batch = synth.stage2_inputs(Bgen, cfg.num_cams, cfg.num_cam_tokens, cfg.num_cond_tokens, cfg.vocab_size, cfg.cond_vocab_size, seed=seed_in)
    return cfg, sd, cam[:B].contiguous(), bev[:B].contiguous(), {k: v[:B].contiguous() for k, v in batch.items()}


def golden_layouts(g):
    import numpy as np
    import torch
    shape = tuple(int(v) for v in g["layout_shape"])
    n = int(np.prod(shape))
    return torch.from_numpy(np.unpackbits(g["layout_bits"])[:n].reshape(shape).astype(np.int64))
VQGAN_CASES = {
    "small_rgb": (dict(in_channels=3, ch=64), 2, 64, 64),
    "small_bev": (dict(in_channels=7, ch=64), 1, 64, 64),
    "config1_rgb": (dict(in_channels=3, ch=128), 2, 128, 128),
}


def gpt_sizes(cfg):
    return dict(num_embed=cfg.num_embed, gpt_block_size=cfg.gpt_block_size, num_img_tokens=cfg.num_img_tokens,
                num_cond_tokens=cfg.num_cond_tokens, num_cams=cfg.num_cams, vocab_size=cfg.vocab_size,
                cond_vocab_size=cfg.cond_vocab_size, num_layers=cfg.num_layers)
