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
    geo = gpt_oracle.geo_from_config(model.cfg)
    sd_t = {k: v.cpu() for k, v in model.transformer.state_dict().items()}
    with torch.no_grad():
        want = gpt_oracle.forward(sd_t, geo, target.cpu().view(1, 6, 256).clone(), ci.view(1, -1), batch, sampling=False)
    if agree == 1.0 and torch.equal(ci.view(1, -1), model.encode_to_c(c.cuda(), batch)[1].cpu()):
        assert (logits.cpu() - want).abs().max().item() < 1e-3
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
