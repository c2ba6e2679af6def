"""the MaskGit stage-2 variant on the CUDA library against goldens minted from the unmodified reference
(modules/stage2/muse_maskgit_pytorch.py) and against the CPU oracle."""
import numpy as np
import pytest
import torch

from bevgen_b200.gpt_config import GPTConfig
from oracle import gpt_oracle, maskgit_oracle, synth
from tests.cases import GPT_SMALL, gpt_sizes

pytestmark = pytest.mark.gpu
LOGIT_TOL = 1e-3          # north_star's floating-point bar


def _case(B=1, precision="fp32x3"):
    from bevgen_b200.maskgit_engine import MaskGitEngine
    cfg = GPTConfig(**GPT_SMALL)
    sd = synth.maskgit_state_dict(gpt_sizes(cfg), 2, cfg.num_heads, seed=3)
    critic = {"weight": sd.pop("to_pred.weight"), "bias": sd.pop("to_pred.bias")}
    cam, bev, batch = synth.stage2_inputs(B, cfg.num_cams, cfg.num_cam_tokens, cfg.num_cond_tokens, cfg.vocab_size, cfg.cond_vocab_size, seed=6)
    eng = MaskGitEngine(sd, cfg, depth=2, heads=cfg.num_heads, device="cuda:0", precision=precision, critic=critic)
    sd_all = dict(sd, **{"to_pred.weight": critic["weight"], "to_pred.bias": critic["bias"]})
    return cfg, sd_all, cam, bev, batch, eng


@pytest.mark.parametrize("precision", ["fp32x3", "f16f8"])
def test_forward_vs_reference_golden(precision, golden_dir):
    g = np.load(golden_dir / "maskgit_small.npz")
    cfg, sd, cam, bev, batch, eng = _case(precision=precision)
    ids = torch.from_numpy(g["ids"]).long()
    logits, emb = eng.forward(ids.cuda(), bev.cuda(), batch)
    torch.cuda.synchronize()
    cols = g["cols"]
    e_l = (logits[:, cols].cpu() - torch.from_numpy(g["logits"])).abs().max().item()
    e_e = (emb[:, cols].cpu() - torch.from_numpy(g["embed"])).abs().max().item()
    print(f"[maskgit small {precision}] logits err {e_l:.2e} (absmax {float(g['logits_absmax']):.2f}), embed err {e_e:.2e}")
    assert torch.isfinite(logits).all()
    assert e_l < LOGIT_TOL and e_e < LOGIT_TOL
    assert abs(logits.double().mean().item() - float(g["logits_mean"])) < 1e-4


def test_forward_batch2_and_critic_vs_oracle():
    cfg, sd, cam, bev, batch, eng = _case(B=2)
    ids = cam.reshape(2 * cfg.num_cams, cfg.num_cam_tokens).clone()
    gen = torch.Generator().manual_seed(9)
    ids[torch.rand(ids.shape, generator=gen) < 0.5] = cfg.vocab_size
    geo = gpt_oracle.geo_from_config(cfg)
    with torch.no_grad():
        want_l, want_e = maskgit_oracle.forward(sd, geo, ids, bev, batch, 2, cfg.num_heads)
        want_s = (want_e @ sd["to_pred.weight"].t() + sd["to_pred.bias"])[..., 0]
    logits, emb = eng.forward(ids.cuda(), bev.cuda(), batch)
    scores = eng.critic_scores(ids.cuda(), bev.cuda(), batch)
    torch.cuda.synchronize()
    assert (logits.cpu() - want_l).abs().max().item() < LOGIT_TOL
    assert (emb.cpu() - want_e).abs().max().item() < LOGIT_TOL
    assert (scores.cpu() - want_s).abs().max().item() < LOGIT_TOL


def test_generate_replays_the_reference_draws(golden_dir):
    """6-step generate with the SelfCritic, fed the same uniform draws as the reference's seeded CPU run: every de-masking step's logits
    match the oracle's on the oracle's own ids, and the final token ids equal the reference's golden ids."""
    g = np.load(golden_dir / "maskgit_small.npz")
    cfg, sd, cam, bev, batch, eng = _case()
    geo = gpt_oracle.geo_from_config(cfg)
    steps = int(g["gen_steps"])
    draws = []

    def rec_noise(kind, step, shape):
        u = torch.zeros(shape).float().uniform_(0, 1)
        draws.append(u)
        return u
    torch.manual_seed(int(g["gen_seed"]))
    otrace = []
    with torch.no_grad():
        want = maskgit_oracle.generate(sd, geo, bev, batch, 2, cfg.num_heads, rec_noise, timesteps=steps, trace=otrace)
    assert np.array_equal(want.numpy().astype(np.int32), g["generated"])
    it = iter(draws)
    trace = []
    got = eng.generate(bev, batch, timesteps=steps, noise=lambda kind, step, shape: next(it), trace=trace)
    torch.cuda.synchronize()
    for (oi, ol), (gi, gl) in zip(otrace, trace):
        if torch.equal(oi, gi.cpu()):                      # same input ids -> logits within the bar
            assert (ol - gl.cpu()).abs().max().item() < LOGIT_TOL
    same = (got.cpu() == want).float().mean().item()
    print(f"[maskgit generate] identical tokens {same:.4f}")
    assert torch.equal(trace[0][0].cpu(), otrace[0][0])
    assert same == 1.0
    assert int(got.max()) < cfg.vocab_size


def test_module_under_the_reference_import_path(golden_dir):
    """multi_view_generation.modules.stage2.muse_maskgit_pytorch: reference constructor arguments and state_dict keys (every key of the
    reference's transformer loads, only its non-trainable buffers are absent from the synthetic dict), forward / generate on the engine."""
    from multi_view_generation.modules.stage2 import muse_maskgit_pytorch as m
    g = np.load(golden_dir / "maskgit_small.npz")
    cfg = GPTConfig(**GPT_SMALL)
    tr = m.MaskGitTransformerMultiView(num_tokens=cfg.vocab_size, dim=cfg.num_embed, seq_len=tuple(cfg.cam_latent_res), depth=2, dim_head=64,
                                       heads=cfg.num_heads, ff_mult=4, cfg=cfg)
    mg = m.MaskGit(image_size=tuple(cfg.cam_latent_res), transformer=tr, self_token_critic=True).eval()
    sd = synth.maskgit_state_dict(gpt_sizes(cfg), 2, cfg.num_heads, seed=3)
    crit = {k[len("to_pred."):]: sd.pop(k) for k in ("to_pred.weight", "to_pred.bias")}
    missing, unexpected = tr.load_state_dict(sd, strict=False)
    assert not unexpected and all(k.endswith(".beta") or k == "bev_grid" for k in missing)
    mg.token_critic.to_pred.load_state_dict(crit)
    with pytest.raises(RuntimeError):
        tr(torch.from_numpy(g["ids"]).long(), conditioning_token_ids=torch.zeros(1, 256, dtype=torch.long), batch={})     # no CPU fallback
    mg = mg.cuda()
    cam, bev, batch = synth.stage2_inputs(1, cfg.num_cams, cfg.num_cam_tokens, cfg.num_cond_tokens, cfg.vocab_size, cfg.cond_vocab_size, seed=6)
    ids = torch.from_numpy(g["ids"]).long().cuda()
    logits, emb = tr(ids, return_embed=True, conditioning_token_ids=bev.cuda(), batch=batch)
    guided = tr.forward_with_cond_scale(ids, conditioning_token_ids=bev.cuda(), batch=batch, cond_scale=3.0)
    assert torch.equal(guided, logits)
    assert (logits[:, g["cols"]].cpu() - torch.from_numpy(g["logits"])).abs().max().item() < LOGIT_TOL
    mg.sample_seed = 7
    a = mg.generate(cond_images=bev.cuda(), fmap_size=tuple(cfg.cam_latent_res), batch=batch, timesteps=4)
    b = mg.generate(cond_images=bev.cuda(), fmap_size=tuple(cfg.cam_latent_res), batch=batch, timesteps=4)
    assert a.shape == (cfg.num_cams, 16, 16) and torch.equal(a, b) and int(a.max()) < cfg.vocab_size and int(a.min()) >= 0
    # partial decoding: given cameras keep their tokens
    init = torch.full((cfg.num_cams, cfg.num_cam_tokens), mg.mask_id, dtype=torch.long, device="cuda")
    init[2] = cam.reshape(cfg.num_cams, -1)[2].cuda()
    c = mg.generate(init_ids=init, cond_images=bev.cuda(), fmap_size=tuple(cfg.cam_latent_res), batch=batch, timesteps=4)
    assert torch.equal(c.reshape(cfg.num_cams, -1)[2], init[2])


@pytest.mark.parametrize("precision", ["fp32x3", "f16f8"])
def test_small_rig_with_fewer_image_than_context_tokens(precision):
    """3 cameras x 7x9 latents (189 image tokens, 256 context tokens, sparse_block_size 1 as in the reference's MaskGit config): the shared
    q | k | v plane is sized by the cross-attention keys, self-attention skips the key tiles beyond its own; checked against the oracle."""
    from bevgen_b200.maskgit_engine import MaskGitEngine
    kw = {**GPT_SMALL, "num_cams": 3, "cam_names": "NUSCENES_ABLATION_CAMERAS", "cam_latent_res": (7, 9), "cam_res": (112, 144), "sparse_block_size": 1}
    cfg = GPTConfig(**kw)
    assert cfg.num_pad_tokens == 0
    sd = synth.maskgit_state_dict(gpt_sizes(cfg), 2, cfg.num_heads, seed=8)
    critic = {"weight": sd.pop("to_pred.weight"), "bias": sd.pop("to_pred.bias")}
    B = 2
    cam, bev, batch = synth.stage2_inputs(B, cfg.num_cams, cfg.num_cam_tokens, cfg.num_cond_tokens, cfg.vocab_size, cfg.cond_vocab_size, seed=2)
    ids = cam.reshape(B * cfg.num_cams, cfg.num_cam_tokens).clone()
    ids[:, ::3] = cfg.vocab_size
    eng = MaskGitEngine(sd, cfg, depth=2, heads=cfg.num_heads, device="cuda:0", precision=precision, critic=critic)
    assert eng.lk_f == 384 and eng.lk_self == 256 and eng.self_tiles is not None
    with torch.no_grad():
        want_l, want_e = maskgit_oracle.forward(sd, gpt_oracle.geo_from_config(cfg), ids, bev, batch, 2, cfg.num_heads)
    logits, emb = eng.forward(ids.cuda(), bev.cuda(), batch)
    torch.cuda.synchronize()
    assert (logits.cpu() - want_l).abs().max().item() < LOGIT_TOL
    assert (emb.cpu() - want_e).abs().max().item() < LOGIT_TOL
    eng.fused_self = False                        # composed Q.K^T -> softmax -> P.V path agrees
    l2, _ = eng.forward(ids.cuda(), bev.cuda(), batch)
    assert (l2 - logits).abs().max().item() < 2e-4


def test_net2net_muse_wrapper_generates_images():
    """cond_transformer_multi_view_muse.Net2NetTransformer (reference :28-286): sample / log_images with both VQGANs and the MaskGit decoder
    under the reference's import paths; the given camera of a partial decoding keeps its ground-truth tokens."""
    from multi_view_generation.modules.losses.vqperceptual import DummyLoss
    from multi_view_generation.modules.stage1.vqgan import VQModel, VQSegmentationModel
    from multi_view_generation.modules.stage2 import muse_maskgit_pytorch as m
    from multi_view_generation.modules.stage2.cond_transformer_multi_view_muse import Net2NetTransformer
    cfg = GPTConfig(**{**GPT_SMALL, "vocab_size": 1024, "cond_vocab_size": 1024})
    tr = m.MaskGitTransformerMultiView(num_tokens=cfg.vocab_size, dim=cfg.num_embed, seq_len=tuple(cfg.cam_latent_res), depth=2, dim_head=64,
                                       heads=cfg.num_heads, ff_mult=4, cfg=cfg)
    mg = m.MaskGit(image_size=tuple(cfg.cam_latent_res), transformer=tr, self_token_critic=True)
    sd = synth.maskgit_state_dict(gpt_sizes(cfg), 2, cfg.num_heads, seed=3)
    mg.token_critic.to_pred.load_state_dict({"weight": sd.pop("to_pred.weight"), "bias": sd.pop("to_pred.bias")})
    tr.load_state_dict(sd, strict=False)
    dd, ddb = synth.vqgan_ddconfig(ch=64), synth.vqgan_ddconfig(ch=64, in_channels=7)
    fs = VQModel(dd, DummyLoss(), 1024, 256, (256, 256), (16, 16), 256)
    fs.load_state_dict(synth.vqgan_state_dict(dd, seed=1))
    cs = VQSegmentationModel(7, ddb, DummyLoss(), 1024, 256, (256, 256), (16, 16), 256)
    cs.load_state_dict(synth.vqgan_state_dict(ddb, seed=5), strict=False)
    model = Net2NetTransformer(mg, fs, cs, cfg, sample_iterations=4).cuda().eval()
    mg.sample_seed = 3
    g = torch.Generator().manual_seed(0)
    batch = {"image": torch.randn(1, 6, 256, 256, 3, generator=g), "segmentation": (torch.rand(1, 256, 256, 7, generator=g) > 0.5).float(),
             "intrinsics_inv": torch.randn(1, 6, 3, 3, generator=g), "extrinsics_inv": torch.randn(1, 6, 4, 4, generator=g)}
    out = model.test_step(batch, 0)
    for k in ("gen", "rec", "gt"):
        assert out[k].shape == (1, 6, 3, 256, 256)
        assert float(out[k].min()) >= 0.0 and float(out[k].max()) <= 1.0
    x, c = model.get_xc(batch)
    _, c_idx = model.encode_to_c(c.cuda(), batch)
    _, z_idx = model.encode_to_z(x.cuda(), batch)
    ids = model.sample(c_idx, batch, partial_decoding_idx=[1, 4])
    assert ids.shape == (6, 16, 16) and int(ids.max()) < 1024
    assert torch.equal(ids.reshape(6, -1)[[1, 4]], z_idx.reshape(6, -1)[[1, 4]])


def test_reference_geometry_14x25_vs_oracle():
    """The geometry of the reference's MaskGit config (6 cameras x 14x25 latents = 2100 image tokens, sparse_block_size 1): ragged 128-row
    tiles everywhere (2100 queries, 2101 / 257 keys in 2176-row planes), checked against the oracle at reduced width."""
    from bevgen_b200.maskgit_engine import MaskGitEngine
    kw = {**GPT_SMALL, "cam_latent_res": (14, 25), "cam_res": (224, 400), "sparse_block_size": 1}
    cfg = GPTConfig(**kw)
    assert cfg.num_img_tokens == 2100 and cfg.num_pad_tokens == 0
    sd = synth.maskgit_state_dict(gpt_sizes(cfg), 2, cfg.num_heads, seed=11)
    critic = {"weight": sd.pop("to_pred.weight"), "bias": sd.pop("to_pred.bias")}
    cam, bev, batch = synth.stage2_inputs(1, cfg.num_cams, cfg.num_cam_tokens, cfg.num_cond_tokens, cfg.vocab_size, cfg.cond_vocab_size, seed=3)
    ids = cam.reshape(cfg.num_cams, cfg.num_cam_tokens).clone()
    ids[:, 1::2] = cfg.vocab_size
    eng = MaskGitEngine(sd, cfg, depth=2, heads=cfg.num_heads, device="cuda:0", precision="f16f8", critic=critic)
    assert eng.lk_f == 2176
    with torch.no_grad():
        want_l, want_e = maskgit_oracle.forward(sd, gpt_oracle.geo_from_config(cfg), ids, bev, batch, 2, cfg.num_heads)
    logits, emb = eng.forward(ids.cuda(), bev.cuda(), batch)
    torch.cuda.synchronize()
    assert torch.isfinite(logits).all()
    assert (logits.cpu() - want_l).abs().max().item() < LOGIT_TOL
    assert (emb.cpu() - want_e).abs().max().item() < LOGIT_TOL


def test_token_bookkeeping_kernels_match_the_torch_statement():
    """bevgen_mg_sample / bevgen_mg_remask against the torch lines of MaskGit.generate they replace (muse_maskgit_pytorch.py:569-619):
    top-k filter + gumbel arg-max + fill of masked positions + `1 - softmax[pred]` scores; top-n re-masking with critic noise and init_ids."""
    from bevgen_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(5)
    R, hw, V, mask_id = 12, 350, 1025, 1024
    logits = torch.randn(R * hw, V, device="cuda", generator=g) * 2.0
    u = torch.rand(R * hw, V, device="cuda", generator=g)
    ids = torch.randint(0, 1024, (R, hw), device="cuda", generator=g)
    ids[torch.rand(R, hw, device="cuda", generator=g) < 0.6] = mask_id
    for temp in (0.7, 0.0):
        k = 103
        gum = -torch.log((-torch.log(u.clamp(min=1e-20))).clamp(min=1e-20))
        val, ind = logits.topk(k, dim=-1)
        filt = torch.full_like(logits, float("-inf")).scatter_(1, ind, val)
        pred = (filt / max(temp, 1e-10) + gum).argmax(-1).view(R, hw)
        is_mask = ids == mask_id
        want_ids = torch.where(is_mask, pred, ids)
        want_sc = (1 - logits.softmax(-1).gather(1, pred.view(-1, 1))[:, 0]).view(R, hw).masked_fill(~is_mask, -1e5)
        got_ids, got_sc = ids.clone(), torch.empty(R, hw, device="cuda")
        ops.mg_sample(logits, u, got_ids, k, 1.0 / max(temp, 1e-10), mask_id, scores=got_sc)
        assert (got_ids != want_ids).float().mean().item() < 2e-4            # a 1-ulp difference of logf may flip a near-tie
        same = got_ids == want_ids
        assert (got_sc - want_sc)[same].abs().max().item() < 1e-5
    # re-masking: exact top-n by rank, critic noise folded in, init_ids restored
    sc = torch.randn(R, hw, device="cuda", generator=g)
    un = torch.rand(R, hw, device="cuda", generator=g)
    init = torch.full((R, hw), mask_id, device="cuda")
    init[:, :40] = 7
    for n_mask in (1, 123, hw):
        eff = sc + (un - 0.5) * 0.8
        want = want_ids.scatter(1, eff.topk(n_mask, dim=-1).indices, mask_id)
        want[init != mask_id] = init[init != mask_id]
        got = want_ids.clone()
        ops.mg_remask(sc, got, n_mask, mask_id, uniform=un, noise_scale=0.8, init_ids=init)
        assert torch.equal(got, want), n_mask
