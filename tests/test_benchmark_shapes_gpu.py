"""GPU parity AT BASELINE.json's OWN SHAPES (the sizes bench.py times), against goldens minted from the unmodified reference:
  configs[1]  ch=128 VQGAN, 96 images 256x256 in one batch (332 tiles per persistent CTA - ring wrap / mbarrier phase territory)
  configs[2]  24-layer, d=1024, 16-head GPT at B=16 (teacher-forced forward)
  configs[3]  KV-cache sampler at B=16 on the 24-layer model (all 1536 steps replayed against the forward)
plus the SURVEY 8f-2 / 8f-4 variants pinned to reference-minted goldens (reference-drawn layouts of density < 1, 3-camera Argoverse
rig, nuScenes-native 14x25 latents).  Reference: modules/stage1/model.py:406-433,506-537, modules/transformer/mingpt_sparse.py:319-391,
modules/stage2/cond_transformer_multi_view.py:154-227."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from bevgen_b200.gpt_config import GPTConfig  # noqa: E402
from bevgen_b200.gpt_decode import GPTSampler  # noqa: E402
from bevgen_b200.gpt_engine import GPTEngine  # noqa: E402
from bevgen_b200.vqgan_engine import VQGANEngine  # noqa: E402
from oracle import synth  # noqa: E402
from tests.cases import GPT_VARIANTS, golden_layouts, gpt_variant_inputs  # noqa: E402

TOL = 1e-3


# ----------------------------------------------------------------------------------------------- configs[1]
@pytest.fixture(scope="module")
def vqgan96():
    dd = synth.vqgan_ddconfig(in_channels=3, ch=128)
    sd = synth.vqgan_state_dict(dd, seed=1)
    x = synth.image_batch(96, 3, 256, 256, seed=100)
    eng = VQGANEngine(sd, dd, device="cuda:0", precision="f16f8")
    xd = x.cuda()
    zq, idx, h = eng.encode(xd)
    rec = eng.decode_indices(idx, 96, 16, 16)
    torch.cuda.synchronize()
    return eng, xd, idx.clone(), eng.nhwc_to_nchw(h).clone(), rec.clone()


def test_vqgan_96x256x256_matches_reference_golden(vqgan96, golden_dir):
    """The benchmark batch itself: its first 8 images against the reference's latents / tokens / reconstruction."""
    g = np.load(golden_dir / "vqgan_config2_rgb.npz")
    eng, xd, idx, h, rec = vqgan96
    err_h = np.abs(h[:8, ::8].cpu().numpy() - g["h_sub"]).max()
    got_idx = idx[: 8 * 256].cpu().numpy()
    mism = np.nonzero(got_idx != g["idx"])[0]
    assert err_h < TOL
    assert all(g["gap"][r] < 64 * err_h for r in mism), "token mismatch that is not a near tie of the reference"
    assert len(mism) <= 2
    # the decoder alone, from the REFERENCE's tokens
    rec_ref = eng.decode_indices(torch.from_numpy(g["idx"]).long().cuda(), 8, 16, 16)
    err_r = max(np.abs(rec_ref[:, :, ::8, ::8].cpu().numpy() - g["rec_sub"]).max(), np.abs(rec_ref[:, :, 100].cpu().numpy() - g["rec_row"]).max())
    print(f"[config2_rgb 96x256x256] latent err {err_h:.2e}, rec err {err_r:.2e}, token mismatches {len(mism)}/2048 (min reference gap {float(g['min_gap']):.1e})")
    assert err_r < TOL
    assert abs(rec_ref.double().mean().item() - float(g["rec_mean"])) < 1e-4


def test_vqgan_96_batch_equals_12_runs_of_8(vqgan96):
    """Scenes are independent (GroupNorm is per image): one 96-image launch must equal 12 launches of 8 BIT-EXACTLY."""
    eng, xd, idx, h, rec = vqgan96
    for i in range(12):
        zq8, idx8, h8 = eng.encode(xd[8 * i: 8 * i + 8].contiguous())
        rec8 = eng.decode_indices(idx8, 8, 16, 16)
        assert torch.equal(idx8, idx[2048 * i: 2048 * (i + 1)]), f"tokens of images {8 * i}.. differ between batch sizes"
        assert torch.equal(eng.nhwc_to_nchw(h8), h[8 * i: 8 * i + 8]), f"latents of images {8 * i}.. differ between batch sizes"
        assert torch.equal(rec8, rec[8 * i: 8 * i + 8]), f"reconstructions of images {8 * i}.. differ between batch sizes"


def test_bev_tokenizer_256x256_matches_reference_golden(golden_dir):
    """The 7-channel BEV tokenizer (VQSegmentationModel ddconfig, configs/model/stage_2_argoverse.yaml:15-18) at 256x256."""
    g = np.load(golden_dir / "vqgan_config2_bev.npz")
    dd = synth.vqgan_ddconfig(in_channels=7, ch=128)
    sd = synth.vqgan_state_dict(dd, seed=1)
    x = (synth.image_batch(16, 7, 256, 256, seed=300)[:2] > 0).float().contiguous()
    eng = VQGANEngine(sd, dd, device="cuda:0", precision="f16f8")
    zq, idx, h = eng.encode(x.cuda())
    err_h = np.abs(eng.nhwc_to_nchw(h)[:, ::8].cpu().numpy() - g["h_sub"]).max()
    mism = np.nonzero(idx.cpu().numpy() != g["idx"])[0]
    rec = eng.decode_indices(torch.from_numpy(g["idx"]).long().cuda(), 2, 16, 16)
    err_r = np.abs(rec[:, :, ::8, ::8].cpu().numpy() - g["rec_sub"]).max()
    print(f"[config2_bev] latent err {err_h:.2e}, rec err {err_r:.2e}, token mismatches {len(mism)}/512")
    assert err_h < TOL and err_r < TOL
    assert all(g["gap"][r] < 64 * err_h for r in mism) and len(mism) <= 1


# ----------------------------------------------------------------------------------------------- configs[2] / [3]
@pytest.fixture(scope="module")
def gpt24():
    cfg, sd, cam, bev, batch = gpt_variant_inputs("full24", synth, GPTConfig, B=16)
    eng = GPTEngine(sd, cfg, device="cuda:0", precision="f16f8")
    del sd
    return cfg, cam, bev, batch, eng


def test_gpt_24_layers_B16_matches_reference_golden(gpt24, golden_dir):
    """The benchmark model and batch: samples 0..1 of the B=16 forward against the reference's logits (24 layers, sampled rows); every
    other sample against a B=2 launch of the same engine."""
    g = np.load(golden_dir / "gpt_full24.npz")
    cfg, cam, bev, batch, eng = gpt24
    rows = g["rows"]
    bd = {k: v.cuda() for k, v in batch.items()}
    tf, hid = eng.forward(cam.clone().cuda(), bev.cuda(), bd, sampling=False, return_hidden=True)
    e_tf = np.abs(tf[:2, rows].cpu().numpy() - g["logits_tf"]).max()
    e_h0 = np.abs(hid[0][:2, ::97].cpu().numpy() - g["hidden0_rows"]).max()
    e_hl = np.abs(hid[-1][:2, ::97].cpu().numpy() - g["hidden_last_rows"]).max()
    del hid
    s = eng.forward(cam.cuda(), bev.cuda(), bd, sampling=True)
    e_s = np.abs(s[:2, rows].cpu().numpy() - g["logits_s"]).max()
    print(f"[full24 B=16] logits tf {e_tf:.2e} sampling {e_s:.2e} hidden0 {e_h0:.2e} hidden_last {e_hl:.2e} (|logit| max {float(g['logits_tf_absmax']):.2f})")
    assert max(e_tf, e_s, e_h0) < TOL and e_hl < 2 * TOL        # hidden_last is the un-normalised residual stream (|x| ~ 10)
    assert abs(tf[:2].double().mean().item() - float(g["logits_tf_mean"])) < 1e-5
    for b0 in (6, 14):
        sub = eng.forward(cam[b0: b0 + 2].cuda(), bev[b0: b0 + 2].cuda(), {k: v[b0: b0 + 2].contiguous() for k, v in bd.items()}, sampling=True)
        assert (sub - s[b0: b0 + 2]).abs().max().item() < 1e-5, "a sample's logits depend on the batch it is launched in"


def test_kv_cache_sampler_B16_24_layers_replays_the_forward(gpt24, golden_dir):
    """configs[3] shape: 16 scenes, 24 layers, all 1536 cached steps with the ground-truth tokens forced; the traced logits must equal the
    reference golden rows (samples 0..1) and the engine's teacher-forced forward (all 16 samples, all positions)."""
    g = np.load(golden_dir / "gpt_full24.npz")
    cfg, cam, bev, batch, eng = gpt24
    B = 16
    bd = {k: v.cuda() for k, v in batch.items()}
    forced = cam.reshape(B, -1)[:, cfg.forward_shuffle_idx]
    sampler = GPTSampler(eng, B)
    toks, trace = sampler.sample(bev, bd, forced_tokens=forced, trace_logits=True)
    torch.cuda.synchronize()
    assert torch.equal(toks.cpu(), cam)
    rows = g["rows"]
    steps = cfg.backward_shuffle_idx[rows]
    err = np.abs(trace[steps][:, :2].permute(1, 0, 2).cpu().numpy() - g["logits_s"]).max()
    full = eng.forward(cam.cuda(), bev.cuda(), bd, sampling=True)
    want = full[:, cfg.forward_shuffle_idx.cuda()].permute(1, 0, 2)
    err_all = (trace - want).abs().max().item()
    print(f"[full24 B=16] KV-cache replay: max logit err vs reference golden {err:.2e}; vs own forward over all 16 x 1536 rows {err_all:.2e}")
    assert err < TOL and err_all < TOL


# ----------------------------------------------------------------------------------------------- SURVEY 8f-2 / 8f-4
@pytest.mark.parametrize("precision", ["fp32x3", "f16f8"])
@pytest.mark.parametrize("name", ["small_density25", "small_density50", "small_argo3", "small_nusc14x25"])
def test_variants_forward_vs_reference_golden(name, precision, golden_dir):
    g = np.load(golden_dir / f"gpt_{name}.npz")
    cfg, sd, cam, bev, batch = gpt_variant_inputs(name, synth, GPTConfig)
    layouts = golden_layouts(g) if GPT_VARIANTS[name][4] else None
    eng = GPTEngine(sd, cfg, device="cuda:0", precision=precision, layouts=layouts)
    if layouts is not None:
        assert eng.layouts is not None and eng.layers[0]["layout64"] is not None        # fused kernel with the layout bit table
    if name == "small_nusc14x25":
        assert eng.Lrun == 2432 and eng.fused_pad                                        # L = 2368 runs on the fused kernel, S never in HBM
    rows = g["rows"]
    tf, hid = eng.forward(cam.clone().cuda(), bev.cuda(), batch, sampling=False, return_hidden=True)
    s = eng.forward(cam.cuda(), bev.cuda(), batch, sampling=True)
    e_tf = np.abs(tf[:, rows].cpu().numpy() - g["logits_tf"]).max()
    e_s = np.abs(s[:, rows].cpu().numpy() - g["logits_s"]).max()
    e_hl = np.abs(hid[-1][:, ::97].cpu().numpy() - g["hidden_last_rows"]).max()
    print(f"[{name}] {precision}: logits tf {e_tf:.2e} sampling {e_s:.2e} hidden_last {e_hl:.2e}")
    assert max(e_tf, e_s, e_hl) < TOL
    # the composed path (scores -> masked softmax -> P.V) agrees with the fused one
    eng.fused_attention = False
    comp = eng.forward(cam.cuda(), bev.cuda(), batch, sampling=True)
    assert (comp - s).abs().max().item() < 2e-4


@pytest.mark.parametrize("name", ["small_density25", "small_argo3", "small_nusc14x25"])
def test_variants_kv_cache_decode_vs_reference_golden(name, golden_dir):
    """The KV-cache sampler on the same variants: every cached step replayed against the reference's logits rows."""
    g = np.load(golden_dir / f"gpt_{name}.npz")
    cfg, sd, cam, bev, batch = gpt_variant_inputs(name, synth, GPTConfig)
    layouts = golden_layouts(g) if GPT_VARIANTS[name][4] else None
    eng = GPTEngine(sd, cfg, device="cuda:0", precision="fp32x3", layouts=layouts)
    B = cam.shape[0]
    forced = cam.reshape(B, -1)[:, cfg.forward_shuffle_idx]
    toks, trace = GPTSampler(eng, B).sample(bev, batch, forced_tokens=forced, trace_logits=True)
    torch.cuda.synchronize()
    assert torch.equal(toks.cpu(), cam)
    steps = cfg.backward_shuffle_idx[g["rows"]]
    err = np.abs(trace[steps].permute(1, 0, 2).cpu().numpy() - g["logits_s"]).max()
    print(f"[{name}] KV-cache replay vs reference golden: {err:.2e}")
    assert err < TOL
