"""Drop-in surface of stage 2 (Net2NetTransformer / GPT / VQModel under the reference's import paths) on the GPU."""
import pytest
import torch

pytestmark = pytest.mark.gpu

from multi_view_generation.modules.losses.vqperceptual import DummyLoss  # noqa: E402
from multi_view_generation.modules.stage1.vqgan import VQModel, VQSegmentationModel  # noqa: E402
from multi_view_generation.modules.stage2.cond_transformer_multi_view import Net2NetTransformer  # noqa: E402
from multi_view_generation.modules.transformer.mingpt_sparse import GPT, GPTConfig  # noqa: E402
from oracle import gpt_oracle, synth, vqgan_oracle  # noqa: E402
from tests.cases import GPT_SMALL, gpt_sizes  # noqa: E402


@pytest.fixture(scope="module")
def model():
    cfg = GPTConfig(**{**GPT_SMALL, "vocab_size": 1024, "cond_vocab_size": 1024})
    gpt = GPT(cfg)
    gpt.load_state_dict(synth.gpt_state_dict(gpt_sizes(cfg), seed=2), strict=False)
    dd, ddb = synth.vqgan_ddconfig(ch=64), synth.vqgan_ddconfig(ch=64, in_channels=7)
    fs = VQModel(dd, DummyLoss(), 1024, 256, (256, 256), (16, 16), 256)
    fs.load_state_dict(synth.vqgan_state_dict(dd, seed=1))
    cs = VQSegmentationModel(7, ddb, DummyLoss(), 1024, 256, (256, 256), (16, 16), 256)
    cs.load_state_dict(synth.vqgan_state_dict(ddb, seed=5), strict=False)
    m = Net2NetTransformer(gpt, fs, cs, top_k=100).cuda().eval()
    m.sample_seed = 7
    return m


def _batch(B=1):
    g = torch.Generator().manual_seed(0)
    return {"image": torch.randn(B, 6, 256, 256, 3, generator=g), "segmentation": (torch.rand(B, 256, 256, 7, generator=g) > 0.5).float(),
            "intrinsics_inv": torch.randn(B, 6, 3, 3, generator=g), "extrinsics_inv": torch.randn(B, 6, 4, 4, generator=g)}


def test_forward_contract_and_oracle(model):
    batch = _batch()
    x, c = model.get_xc(batch)
    assert x.shape == (6, 3, 256, 256) and c.shape == (1, 7, 256, 256)
    logits, target = model(x.cuda(), c.cuda(), batch)
    assert logits.shape == (1, 1536, 1024) and target.shape == (1, 1536) and target.dtype == torch.int64
    # the same tokens through the CPU oracle
    sd_f = {k: v.cpu() for k, v in model.first_stage_model.state_dict().items()}
    sd_c = {k: v.cpu() for k, v in model.cond_stage_model.state_dict().items()}
    with torch.no_grad():
        _, zi, _ = vqgan_oracle.encode(x, sd_f)
        _, ci, _ = vqgan_oracle.encode(c, sd_c)
    agree = (zi.view(1, -1) == target.cpu()).float().mean().item()
    assert agree > 0.99, agree
    c_ours = model.encode_to_c(c.cuda(), batch)[1].cpu()
    assert (ci.view(1, -1) == c_ours).float().mean().item() > 0.99
    # a18: the transformer leg is checked UNCONDITIONALLY - the oracle is fed the module's own tokens, so a near-tie in the VQ argmin
    # (asserted > 99 % agreement above) cannot hide a logits error
    geo = gpt_oracle.geo_from_config(model.cfg)
    sd_t = {k: v.cpu() for k, v in model.transformer.state_dict().items()}
    with torch.no_grad():
        want = gpt_oracle.forward(sd_t, geo, target.cpu().view(1, 6, 256).clone(), c_ours, batch, sampling=False)
    err = (logits.cpu() - want).abs().max().item()
    assert err < 1e-3, err
    loss = model.shared_step(batch)
    assert torch.isfinite(loss)


def test_sample_and_log_images(model):
    batch = _batch()
    _, c = model.get_xc(batch)
    _, c_idx = model.encode_to_c(c.cuda(), batch)
    toks = model.sample(torch.zeros(1, 0), c_idx, batch, temperature=1.0, sample=True, top_k=100)
    assert toks.shape == (1, 6, 256) and toks.dtype == torch.int64 and int(toks.max()) < 1024 and int(toks.min()) >= 0
    toks2 = model.sample(torch.zeros(1, 0), c_idx, batch, temperature=1.0, sample=True, top_k=100)
    assert torch.equal(toks, toks2)                          # same seed -> same draw
    greedy = model.sample(torch.zeros(1, 0), c_idx, batch, sample=False)
    assert not torch.equal(greedy, toks)
    out = model.test_step(batch, 0)
    for k in ("gen", "rec", "gt"):
        assert out[k].shape == (1, 6, 3, 256, 256)
        assert float(out[k].min()) >= 0.0 and float(out[k].max()) <= 1.0
    assert torch.isfinite(model.last_test_loss)
    # a20: the three image sets against the CPU oracle (reference log_images :489-497,525-527 + bev_utils/util.py:97-118)
    x, _ = model.get_xc(batch)
    sd_f = {k: v.cpu() for k, v in model.first_stage_model.state_dict().items()}
    mean, std = torch.tensor([0.4265, 0.4489, 0.4769]).view(1, 3, 1, 1), torch.tensor([0.2053, 0.2206, 0.2578]).view(1, 3, 1, 1)
    denorm = lambda t: (t * std + mean).clamp(0, 1)
    _, z_idx = model.encode_to_z(x.cuda(), batch)
    with torch.no_grad():
        rec_want = denorm(vqgan_oracle.decode(vqgan_oracle.get_codebook_entry(z_idx.reshape(-1).cpu(), (6, 16, 16, 256), sd_f), sd_f))
        gen_want = denorm(vqgan_oracle.decode(vqgan_oracle.get_codebook_entry(toks.reshape(-1).cpu(), (6, 16, 16, 256), sd_f), sd_f))
    e_rec = (out["rec"][0].cpu() - rec_want).abs().max().item()
    e_gen = (out["gen"][0].cpu() - gen_want).abs().max().item()      # sample_seed is fixed: test_step drew the same tokens as `toks`
    e_gt = (out["gt"][0].cpu() - denorm(x)).abs().max().item()
    print(f"log_images vs oracle: rec {e_rec:.2e} gen {e_gen:.2e} gt {e_gt:.2e}")
    assert e_rec < 1e-3 and e_gen < 1e-3 and e_gt < 1e-6


def test_log_images_partial_decoding_modes(model):
    """log_images draws partial_decoding_idx like the reference (:503-515) and passes it to sample; mode 4 = cameras [3, 0, 2]
    keep their ground-truth tokens, i.e. their generated image is the reconstruction (inside the 3-px green frame)."""
    batch = _batch()
    model.partial_decoding = 4
    try:
        out = model.log_images(batch, generate_only=True, top_k=100)
    finally:
        model.partial_decoding = None
    kept, free = [3, 0, 2], [1, 4, 5]
    inner = (slice(None), slice(None), slice(None), slice(3, -3), slice(3, -3))
    assert torch.equal(out["gen"][:, kept][inner], out["rec"][:, kept][inner])
    assert not torch.equal(out["gen"][:, free][inner], out["rec"][:, free][inner])
    frame = out["gen"][0, 3, :, 0, :]                       # top row of a kept camera: (0, 249/255, 0)
    assert torch.allclose(frame, torch.tensor([0.0, 249.0 / 255.0, 0.0], device=frame.device).view(3, 1).expand_as(frame))


def test_partial_decoding_keeps_given_cameras_and_is_consistent(model):
    """partial_decoding_idx (reference :161-165,181-182): the listed cameras keep their ground-truth tokens, the others are decoded
    conditioned on them.  Consistency: under the causal mask the greedy token at every decoded position must be the argmax of the
    teacher-forced logits of the finished sequence (near-ties excepted)."""
    batch = _batch()
    x, c = model.get_xc(batch)
    _, c_idx = model.encode_to_c(c.cuda(), batch)
    _, z = model.encode_to_z(x.cuda(), batch)
    z = z.view(1, 6, 256)
    given = [0, 3]
    toks = model.sample(torch.zeros(1, 0), c_idx, batch, sample=False, partial_decoding_idx=given)
    assert toks.shape == (1, 6, 256) and int(toks.max()) < 1024 and int(toks.min()) >= 0
    assert torch.equal(toks[:, given], z[:, given])
    free = [i for i in range(6) if i not in given]
    assert not torch.equal(toks[:, free], z[:, free])
    logits = model.transformer(toks.clone(), c_idx, batch, sampling=True).view(1, 6, 256, -1)      # (cam,h,w) order
    top2 = logits.topk(2, dim=-1)
    clear = (top2.values[..., 0] - top2.values[..., 1]) > 1e-3                                     # positions without a near-tie
    agree = (top2.indices[..., 0] == toks)[:, free][clear[:, free]]
    assert agree.numel() > 500 and bool(agree.all())
    # and without partial decoding nothing is pinned
    plain = model.sample(torch.zeros(1, 0), c_idx, batch, sample=False)
    assert not torch.equal(plain[:, given], z[:, given])


def test_no_cpu_fallback():
    cfg = GPTConfig(**GPT_SMALL)
    gpt = GPT(cfg)
    _, bev, batch = synth.stage2_inputs(1, 6, 256, 256, 128, 128)
    with pytest.raises(RuntimeError, match="CUDA device only"):
        gpt(torch.zeros(1, 6, 256, dtype=torch.int64), bev, batch, sampling=True)


def test_model_built_from_yaml_targets_runs_test_step():
    """VERDICT r1 'Hydra _target_ round trip': the model is built from a config of the reference's shape (tests/configs/stage_2_small.yaml:
    nested `_target_` paths as in configs/model/stage_2.yaml) and runs the generate.py hot path."""
    from pathlib import Path
    from multi_view_generation.utils.instantiate import instantiate, load_yaml
    cfg = load_yaml(Path(__file__).parent / "configs" / "stage_2_small.yaml")
    m = instantiate(cfg["model"])
    gpt_cfg = m.cfg
    m.transformer.load_state_dict(synth.gpt_state_dict(gpt_sizes(gpt_cfg), seed=2), strict=False)
    m.first_stage_model.load_state_dict(synth.vqgan_state_dict(cfg["model"]["first_stage"]["ddconfig"], seed=1))
    m.cond_stage_model.load_state_dict(synth.vqgan_state_dict(cfg["model"]["cond_stage"]["ddconfig"], seed=5), strict=False)
    m = m.cuda().eval()
    m.sample_seed = 3
    out = m.test_step(_batch(), 0)
    for k in ("gen", "rec", "gt"):
        assert out[k].shape == (1, 6, 3, 256, 256) and torch.isfinite(out[k]).all()
        assert float(out[k].min()) >= 0.0 and float(out[k].max()) <= 1.0
    assert m.top_k == 20 and torch.isfinite(m.last_test_loss)
