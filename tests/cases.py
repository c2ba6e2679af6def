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
VQGAN_CASES = {
    "small_rgb": (dict(in_channels=3, ch=64), 2, 64, 64),
    "small_bev": (dict(in_channels=7, ch=64), 1, 64, 64),
    "config1_rgb": (dict(in_channels=3, ch=128), 2, 128, 128),
}


def gpt_sizes(cfg):
    return dict(num_embed=cfg.num_embed, gpt_block_size=cfg.gpt_block_size, num_img_tokens=cfg.num_img_tokens,
                num_cond_tokens=cfg.num_cond_tokens, num_cams=cfg.num_cams, vocab_size=cfg.vocab_size,
                cond_vocab_size=cfg.cond_vocab_size, num_layers=cfg.num_layers)
