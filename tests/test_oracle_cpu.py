"""Pins the CPU restatements in oracle/ against goldens minted from the unmodified reference."""
import zlib

import numpy as np
import pytest
import torch

from bevgen_b200.gpt_config import GPTConfig
from oracle import gpt_oracle, synth, vqgan_oracle
from tests.cases import GPT_CASES, GPT_SMALL, GPT_VARIANTS, VQGAN_CASES, golden_layouts, gpt_sizes, gpt_variant_inputs


def crc(t):
    return zlib.crc32(np.ascontiguousarray(t.detach().cpu().numpy()).tobytes())


@pytest.mark.parametrize("cb", ["normal", "default"])
def test_vq_oracle(cb, golden_dir):
    g = np.load(golden_dir / f"vq_{cb}.npz")
    sd = synth.vqgan_state_dict(synth.vqgan_ddconfig(), seed=3, codebook=cb)
    z = synth.tensor_for("vq.z", (4, 256, 8, 8), seed=5, kind="embedding")
    if cb == "default":
        z = z * 1e-3
    assert crc(z) == int(g["z_crc"]), "synthetic input drifted (torch RNG change?)"
    zq, idx, _ = vqgan_oracle.vq_nearest(z, sd["quantize.embedding.weight"])
    assert np.array_equal(idx.numpy(), g["idx"])
    assert crc(zq) == int(g["zq_crc"])
    assert crc(vqgan_oracle.get_codebook_entry(idx, (4, 8, 8, 256), sd)) == int(g["zq_crc"])


@pytest.mark.parametrize("name", list(VQGAN_CASES))
def test_vqgan_oracle(name, golden_dir):
    kw, n, H, W = VQGAN_CASES[name]
    g = np.load(golden_dir / f"vqgan_{name}.npz")
    dd = synth.vqgan_ddconfig(**kw)
    sd = synth.vqgan_state_dict(dd, seed=1)
    x = synth.image_batch(n, dd["in_channels"], H, W, seed=7)
    assert crc(x) == int(g["x_crc"])
    with torch.no_grad():
        quant, idx, h = vqgan_oracle.encode(x, sd)
        rec = vqgan_oracle.decode(vqgan_oracle.get_codebook_entry(idx, (n, H // 16, W // 16, 256), sd), sd)
    assert idx.shape == (n * (H // 16) * (W // 16),) and idx.dtype == torch.int64
    assert rec.shape == (n, dd["out_ch"], H, W)
    np.testing.assert_allclose(h.numpy(), g["h"], rtol=0, atol=2e-5)
    assert np.array_equal(idx.numpy(), g["idx"])
    np.testing.assert_allclose(rec.numpy(), g["rec"], rtol=0, atol=2e-5)


def test_vqgan_oracle_at_benchmark_resolution(golden_dir):
    """BASELINE configs[1] shape (ch=128, 256x256): the first 2 of the golden's 8 images (GroupNorm is per image, so a prefix is exact)."""
    g = np.load(golden_dir / "vqgan_config2_rgb.npz")
    dd = synth.vqgan_ddconfig(in_channels=3, ch=128)
    sd = synth.vqgan_state_dict(dd, seed=1)
    x8 = synth.image_batch(96, 3, 256, 256, seed=100)[:8].contiguous()
    assert crc(x8) == int(g["x_crc"])
    x = x8[:2]
    with torch.no_grad():
        quant, idx, h = vqgan_oracle.encode(x, sd)
        rec = vqgan_oracle.decode(quant, sd)
    np.testing.assert_allclose(h[:, ::8].numpy(), g["h_sub"][:2], rtol=0, atol=2e-5)
    assert np.array_equal(idx.numpy(), g["idx"][:512])
    np.testing.assert_allclose(rec[:, :, ::8, ::8].numpy(), g["rec_sub"][:2], rtol=0, atol=3e-5)
    np.testing.assert_allclose(rec[:, :, 100].numpy(), g["rec_row"][:2], rtol=0, atol=3e-5)


@pytest.mark.parametrize("name", ["small_density25", "small_density50", "small_argo3", "small_nusc14x25"])
def test_gpt_oracle_variants(name, golden_dir):
    """SURVEY 8f-2 / 8f-4 pinned to the reference: layouts of density < 1 drawn by the reference itself, the 3-camera Argoverse rig,
    and the nuScenes-native 14x25 latents (L = 2368, 12 pad tokens)."""
    g = np.load(golden_dir / f"gpt_{name}.npz")
    cfg, sd, cam, bev, batch = gpt_variant_inputs(name, synth, GPTConfig)
    assert crc(cam) == int(g["cam_crc"])
    layouts = golden_layouts(g) if GPT_VARIANTS[name][4] else None
    geo = gpt_oracle.geo_from_config(cfg)
    with torch.no_grad():
        tf, hid = gpt_oracle.forward(sd, geo, cam, bev, batch, sampling=False, return_hidden=True, layouts=layouts)
        s = gpt_oracle.forward(sd, geo, cam, bev, batch, sampling=True, layouts=layouts)
    rows = g["rows"]
    np.testing.assert_allclose(tf[:, rows].numpy(), g["logits_tf"], rtol=0, atol=2e-5)
    np.testing.assert_allclose(s[:, rows].numpy(), g["logits_s"], rtol=0, atol=2e-5)
    np.testing.assert_allclose(hid[-1][:, ::97].numpy(), g["hidden_last_rows"], rtol=0, atol=2e-5)
    if layouts is not None:
        dense = gpt_oracle.forward(sd, geo, cam, bev, batch, sampling=True)
        assert (dense - s).abs().max().item() > 1e-2          # the reference-drawn layouts really remove attended positions


def _gpt_case(name):
    kw, B = GPT_CASES[name]
    cfg = GPTConfig(**kw)
    sd = synth.gpt_state_dict(gpt_sizes(cfg), seed=2)
    cam, bev, batch = synth.stage2_inputs(B, cfg.num_cams, cfg.num_cam_tokens, cfg.num_cond_tokens,
                                          cfg.vocab_size, cfg.cond_vocab_size, seed=4)
    return cfg, sd, cam, bev, batch


@pytest.mark.parametrize("name", ["small", "padded", "wide2"])
def test_gpt_oracle(name, golden_dir):
    g = np.load(golden_dir / f"gpt_{name}.npz")
    cfg, sd, cam, bev, batch = _gpt_case(name)
    assert crc(cam) == int(g["cam_crc"])
    geo = gpt_oracle.geo_from_config(cfg)
    with torch.no_grad():
        tf, hid = gpt_oracle.forward(sd, geo, cam, bev, batch, sampling=False, return_hidden=True)
        s = gpt_oracle.forward(sd, geo, cam, bev, batch, sampling=True)
    rows = g["rows"]
    np.testing.assert_allclose(tf[:, rows].numpy(), g["logits_tf"], rtol=0, atol=2e-5)
    np.testing.assert_allclose(s[:, rows].numpy(), g["logits_s"], rtol=0, atol=2e-5)
    np.testing.assert_allclose(hid[0][:, ::97].numpy(), g["hidden0_rows"], rtol=0, atol=2e-5)
    np.testing.assert_allclose(hid[-1][:, ::97].numpy(), g["hidden_last_rows"], rtol=0, atol=2e-5)
    assert abs(tf.double().mean().item() - float(g["logits_tf_mean"])) < 1e-6


def test_sampling_loop_and_kv_cache_oracle(golden_dir):
    """The reference's own loop (4 greedy steps) is reproduced by both the O(L^2) restatement and the KV-cache one."""
    g = np.load(golden_dir / "gpt_small_sample4.npz")
    cfg, sd, cam, bev, batch = _gpt_case("small")
    geo = gpt_oracle.geo_from_config(cfg)
    with torch.no_grad():
        x1, rows1 = gpt_oracle.sample_reference_loop(sd, geo, bev, batch, steps=4)
        x2, rows2 = gpt_oracle.KVCacheDecoder(sd, geo).run(bev, batch, steps=4)
    np.testing.assert_allclose(rows1.numpy(), g["logits"], rtol=0, atol=2e-5)
    np.testing.assert_allclose(rows2.numpy(), g["logits"], rtol=0, atol=2e-5)
    fwd = cfg.forward_shuffle_idx[:4]
    for x in (x1, x2):
        assert np.array_equal(x.reshape(2, -1)[:, fwd].numpy(), g["tokens"])


def test_kv_cache_equals_full_forward_teacher_forced():
    """Teacher-forced replay: feeding tokens z through the KV-cache decoder gives the rows of the full forward."""
    cfg, sd, cam, bev, batch = _gpt_case("small")
    geo = gpt_oracle.geo_from_config(cfg)
    steps = 40
    forced = cam.reshape(cam.shape[0], -1)[:, cfg.forward_shuffle_idx[:steps]]
    with torch.no_grad():
        full = gpt_oracle.forward(sd, geo, cam, bev, batch, sampling=True)
        _, rows = gpt_oracle.KVCacheDecoder(sd, geo).run(bev, batch, steps=steps, forced_tokens=forced)
    want = full[:, cfg.forward_shuffle_idx[:steps]]
    np.testing.assert_allclose(rows.numpy(), want.numpy(), rtol=0, atol=3e-5)


def test_topk_oracle(golden_dir):
    g = np.load(golden_dir / "topk.npz")
    p = gpt_oracle.sample_probs(torch.from_numpy(g["logits"]), 1.0, int(g["k"]))
    np.testing.assert_allclose(p.numpy(), g["probs"], rtol=0, atol=1e-7)
    assert np.array_equal((p > 0).sum(-1).numpy(), g["kept"])


def _maskgit_case():
    from oracle import maskgit_oracle  # noqa: F401
    cfg = GPTConfig(**GPT_SMALL)
    sd = synth.maskgit_state_dict(gpt_sizes(cfg), 2, cfg.num_heads, seed=3)
    cam, bev, batch = synth.stage2_inputs(1, cfg.num_cams, cfg.num_cam_tokens, cfg.num_cond_tokens, cfg.vocab_size, cfg.cond_vocab_size, seed=6)
    return cfg, sd, bev, batch


def cpu_noise(kind, step, shape):
    """Same draws, in the same order, as the reference's gumbel_noise / uniform helpers on a seeded CPU generator."""
    return torch.zeros(shape).float().uniform_(0, 1)


def test_maskgit_oracle_forward_and_generate(golden_dir):
    """the MaskGit restatement against the unmodified reference (muse_maskgit_pytorch.py) - logits / embeddings of one
    forward on partly masked ids, and the token ids of a 6-step `generate` with the SelfCritic replaying the reference's RNG draws."""
    from oracle import maskgit_oracle
    g = np.load(golden_dir / "maskgit_small.npz")
    cfg, sd, bev, batch = _maskgit_case()
    geo = gpt_oracle.geo_from_config(cfg)
    ids = torch.from_numpy(g["ids"]).long()
    with torch.no_grad():
        logits, emb = maskgit_oracle.forward(sd, geo, ids, bev, batch, 2, cfg.num_heads)
    cols = g["cols"]
    assert (logits[:, cols] - torch.from_numpy(g["logits"])).abs().max().item() < 2e-5
    assert (emb[:, cols] - torch.from_numpy(g["embed"])).abs().max().item() < 2e-5
    assert abs(logits.double().mean().item() - float(g["logits_mean"])) < 1e-6
    torch.manual_seed(int(g["gen_seed"]))
    with torch.no_grad():
        gen = maskgit_oracle.generate(sd, geo, bev, batch, 2, cfg.num_heads, cpu_noise, timesteps=int(g["gen_steps"]))
    assert np.array_equal(gen.numpy().astype(np.int32), g["generated"])


def test_vqgan_geometric_embedding_oracle(golden_dir):
    """VQModel(geometric_embedding=True) - the default of configs/model/stage_1_cam.yaml - minted through the reference's own
    LightningModule class (vqgan.py:62-69,84-121): pre-quantisation latent, token ids and reconstruction."""
    g = np.load(golden_dir / "vqgan_geometric.npz")
    dd = synth.vqgan_ddconfig(in_channels=3, ch=64, resolution=64)
    sd = synth.vqgan_state_dict(dd, seed=4, geometric=True)
    x = synth.image_batch(6, 3, 64, 64, seed=9)
    assert zlib.crc32(x.numpy().tobytes()) == int(g["x_crc"])
    batch = {"intrinsics_inv": torch.from_numpy(g["intrinsics_inv"]), "extrinsics_inv": torch.from_numpy(g["extrinsics_inv"])}
    with torch.no_grad():
        quant, idx, h = vqgan_oracle.encode(x, sd, batch, cam_res=(64, 64))
        rec = vqgan_oracle.decode(quant, sd)
        _, idx_plain, _ = vqgan_oracle.encode(x, sd)
    assert np.abs(h.numpy() - g["h"]).max() < 2e-5
    assert np.array_equal(idx.numpy().astype(np.int32), g["idx"].reshape(-1))
    assert np.abs(rec.numpy() - g["rec"]).max() < 2e-5
    assert not torch.equal(idx, idx_plain)
