"""GPU parity of the stage-2 transformer engine vs goldens minted from the unmodified reference GPT (dense fp32 stand-in for
the absent DeepSpeed ops) and vs the CPU oracle.  Tolerance: north_star's 1e-3 on logits (fp32x3 mode)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from bevgen_b200.gpt_config import GPTConfig  # noqa: E402
from bevgen_b200.gpt_engine import GPTEngine  # noqa: E402
from oracle import gpt_oracle, synth  # noqa: E402
from tests.cases import GPT_CASES, gpt_sizes  # noqa: E402

LOGIT_TOL = 1e-3


def _case(name, precision="fp32x3"):
    kw, B = GPT_CASES[name]
    cfg = GPTConfig(**kw)
    sd = synth.gpt_state_dict(gpt_sizes(cfg), seed=2)
    cam, bev, batch = synth.stage2_inputs(B, cfg.num_cams, cfg.num_cam_tokens, cfg.num_cond_tokens, cfg.vocab_size, cfg.cond_vocab_size, seed=4)
    eng = GPTEngine(sd, cfg, device="cuda:0", precision=precision)
    return cfg, sd, cam, bev, batch, eng


def test_embed_vs_oracle():
    cfg, sd, cam, bev, batch, eng = _case("small")
    geo = gpt_oracle.geo_from_config(cfg)
    for sampling in (True, False):
        want = gpt_oracle.embed(sd, geo, cam, bev, batch, sampling)
        got = eng.embed(cam.cuda(), bev.cuda(), batch, sampling).cpu()
        assert ((got - want).abs() / (1 + want.abs())).max().item() < 2e-6
    cfg, sd, cam, bev, batch, eng = _case("padded")
    want = gpt_oracle.embed(sd, gpt_oracle.geo_from_config(cfg), cam, bev, batch, True)
    assert ((eng.embed(cam.cuda(), bev.cuda(), batch, True).cpu() - want).abs() / (1 + want.abs())).max().item() < 2e-6


@pytest.mark.parametrize("precision", ["fp32x3", "f16f8"])
@pytest.mark.parametrize("name", ["small", "padded", "wide2"])
def test_forward_fp32x3_vs_reference_golden(name, precision, golden_dir):
    """Both parity modes (bf16x3 everywhere / fp16 + 2 x e4m3 MLP GEMMs) against the reference's logits at the 1e-3 bar."""
    g = np.load(golden_dir / f"gpt_{name}.npz")
    cfg, sd, cam, bev, batch, eng = _case(name, precision)
    rows = g["rows"]
    tf, hid = eng.forward(cam.clone().cuda(), bev.cuda(), batch, sampling=False, return_hidden=True)
    s = eng.forward(cam.cuda(), bev.cuda(), batch, sampling=True)
    torch.cuda.synchronize()
    e_h0 = np.abs(hid[0][:, ::97].cpu().numpy() - g["hidden0_rows"]).max()
    e_hl = np.abs(hid[-1][:, ::97].cpu().numpy() - g["hidden_last_rows"]).max()
    e_tf = np.abs(tf[:, rows].cpu().numpy() - g["logits_tf"]).max()
    e_s = np.abs(s[:, rows].cpu().numpy() - g["logits_s"]).max()
    print(f"[{name}] {precision}: hidden0 {e_h0:.2e} hidden_last {e_hl:.2e} logits tf {e_tf:.2e} sampling {e_s:.2e} (|logit| max {float(g['logits_tf_absmax']):.2f})")
    assert e_h0 < LOGIT_TOL and e_hl < LOGIT_TOL and e_tf < LOGIT_TOL and e_s < LOGIT_TOL
    assert abs(tf.double().mean().item() - float(g["logits_tf_mean"])) < 1e-5


def test_forward_bf16_error_budget(golden_dir):
    """Single-pass bf16 (BASELINE config 3 names bf16): cannot meet 1e-3 (naive bf16 3.4e-2); budget documented."""
    g = np.load(golden_dir / "gpt_wide2.npz")
    cfg, sd, cam, bev, batch, eng = _case("wide2", "bf16")
    tf = eng.forward(cam.clone().cuda(), bev.cuda(), batch, sampling=False)
    err = np.abs(tf[:, g["rows"]].cpu().numpy() - g["logits_tf"])
    print(f"[wide2] bf16: logits max err {err.max():.2e} mean {err.mean():.2e}")
    assert err.max() < 6e-2 and err.mean() < 1e-2


@pytest.mark.parametrize("npass", [3, 1])
def test_fused_attention_matches_composed(npass):
    """Fused tcgen05 attention vs the composed (QK^T GEMM -> softmax -> PV GEMM) path on the full-size geometry."""
    from bevgen_b200 import ops
    from tests.cases import GPT_KW
    cfg = GPTConfig(**GPT_KW)
    sd = synth.gpt_state_dict(gpt_sizes(cfg), seed=2)
    eng = GPTEngine(sd, cfg, device="cuda:0", precision="fp32x3" if npass == 3 else "bf16")
    B, L, d = 2, cfg.gpt_block_size, cfg.num_embed
    g = torch.Generator().manual_seed(0)
    qkv = torch.randn(B, L, 3 * d, generator=g).cuda()
    hi, lo = ops.split_planes(qkv, npass)
    y = torch.randn(B, L, d, generator=g).cuda()
    eng.fused_attention = True
    fused = eng.attention((hi, lo), y, B, L)
    eng.fused_attention = False
    comp = eng.attention((hi, lo), y, B, L)
    torch.cuda.synchronize()
    # dense fp64 reference
    q, k, v = [(hi.double() + (lo.double() if lo is not None else 0)).view(B, L, 3, 16, 64)[:, :, i].permute(0, 2, 1, 3) for i in range(3)]
    s = (q @ k.transpose(-1, -2) + eng.bias.double()[None, None]) * 0.125
    s = s.masked_fill(eng.mask_u8[None, None] == 0, float("-inf"))
    ref = y.double() + (torch.softmax(s, -1) @ v).permute(0, 2, 1, 3).reshape(B, L, d)
    e_f, e_c = (fused.double() - ref).abs().max().item(), (comp.double() - ref).abs().max().item()
    print(f"[npass={npass}] fused err {e_f:.2e} composed err {e_c:.2e}")
    assert torch.isfinite(fused).all()
    assert e_f < (3e-4 if npass == 3 else 3e-2)


def _random_layouts(cfg, seed=0, keep=0.5):
    """Per-layer, per-head block layouts that REMOVE allowed positions (density < 1) but keep the diagonal and the first key block,
    so that no row is fully masked (an all-masked row is NaN in the reference too, README.md:113)."""
    nb = cfg.gpt_block_size // cfg.sparse_block_size
    g = torch.Generator().manual_seed(seed)
    lay = torch.rand(cfg.num_layers, cfg.num_heads, nb, nb, generator=g) < keep
    lay |= torch.eye(nb, dtype=torch.bool)[None, None]
    lay[..., 0] = True
    return lay.long()


def test_block_sparse_layouts_forward_and_decode_vs_oracle():
    """per-head block layouts of density < 1 (sparse_self_attention.py:59-60,153-173).  Teacher-forced forward against the
    CPU oracle's dense restatement with the same layouts, and the KV-cache sampler replaying the same tokens against that forward."""
    from bevgen_b200.gpt_decode import GPTSampler
    kw, B = GPT_CASES["small"]
    cfg = GPTConfig(**kw)
    sd = synth.gpt_state_dict(gpt_sizes(cfg), seed=2)
    cam, bev, batch = synth.stage2_inputs(B, cfg.num_cams, cfg.num_cam_tokens, cfg.num_cond_tokens, cfg.vocab_size, cfg.cond_vocab_size, seed=4)
    layouts = _random_layouts(cfg)
    assert not cfg.layout_covers_mask(layouts[0])
    eng = GPTEngine(sd, cfg, device="cuda:0", precision="fp32x3", layouts=layouts)
    assert eng.layouts is not None
    got = eng.forward(cam.cuda(), bev.cuda(), batch, sampling=True)
    dense = GPTEngine(sd, cfg, device="cuda:0", precision="fp32x3").forward(cam.cuda(), bev.cuda(), batch, sampling=True)
    geo = gpt_oracle.geo_from_config(cfg)
    with torch.no_grad():
        want = gpt_oracle.forward({k: v.cpu() for k, v in sd.items()}, geo, cam, bev, batch, sampling=True, layouts=layouts)
    err = (got.cpu() - want).abs().max().item()
    print(f"[small, density<1] logits err vs oracle {err:.2e}; change vs dense {(got - dense).abs().max().item():.2e}")
    assert err < LOGIT_TOL
    assert (got - dense).abs().max().item() > 10 * LOGIT_TOL          # the layouts really removed attended positions
    # the forward above ran the fused kernel with the 16-position layout table (key tiles without layout blocks skipped); the composed
    # path (scores -> masked softmax -> P.V) must agree with it
    assert eng.layers[0]["layout64"] is not None
    eng.fused_attention = False
    composed = eng.forward(cam.cuda(), bev.cuda(), batch, sampling=True)
    eng.fused_attention = True
    assert (got - composed).abs().max().item() < 2e-4
    forced = cam.reshape(B, -1)[:, cfg.forward_shuffle_idx]
    toks, trace = GPTSampler(eng, B).sample(bev, batch, forced_tokens=forced, trace_logits=True, steps=300)
    torch.cuda.synchronize()
    ref_rows = got[:, cfg.forward_shuffle_idx.cuda()[:300]].permute(1, 0, 2)
    assert (trace[:300] - ref_rows).abs().max().item() < 2e-4
